// imma_peak.cu — int8 tensor throughput reachable with warp-level mma.sync (m16n8k32 s8*s8+s32) on B200,
// to decide whether the constant-operand Barrett phases (q1*mu, q*Nt: batch x Toeplitz GEMMs) should move
// off the IMAD pipe.  Reports int8 MAC/s and MAC/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mma_s8(int (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(512) k_imma(int* out, unsigned seed, int iters) {
    int acc[NACC][4];
    unsigned a[4], b[NACC][2];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = seed * (i + 1) + threadIdx.x;
#pragma unroll
    for (int n = 0; n < NACC; n++) { b[n][0] = seed + n; b[n][1] = seed * 3 + n + threadIdx.x; for (int i = 0; i < 4; i++) acc[n][i] = 0; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int n = 0; n < NACC; n++) mma_s8(acc[n], a, b[n]);
            a[u] += (unsigned)acc[0][0];     // loop-carried operand
        }
    }
    int s = 0;
#pragma unroll
    for (int n = 0; n < NACC; n++) for (int i = 0; i < 4; i++) s ^= acc[n][i];
    if (s == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int* out; CK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * 4));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    bool first = true;
    const int iters = 2000;
    for (int threads : {128, 256, 512}) for (int cps : {1, 2, 4}) {
        if (threads * cps > 2048) continue;
        double ms = time_ms([&] { k_imma<8><<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
        double macs = (double)sms * cps * (threads / 32) * iters * 4 * 8 * (16.0 * 8 * 32);
        double per_s = macs / (ms * 1e-3);
        printf("%s  {\"variant\": \"mma_sync_m16n8k32_s8_acc8\", \"threads_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"int8_mac_per_s\": %.6e, \"mac_per_clk_per_sm_at_1965MHz\": %.1f}",
               first ? "" : ",\n", threads, cps, ms, per_s, per_s / sms / 1.965e9);
        first = false; fflush(stdout);
    }
    printf("\n]}\n");
    return 0;
}
