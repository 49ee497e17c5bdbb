// imad_peak2.cu — corrected integer-multiply microbenchmark (supersedes the wide_indep/lo32 variants of
// imad_peak.cu, whose loop-invariant operands let ptxas hoist the multiplies out of the loop).
// Every multiply here has a loop-carried operand; SASS instruction counts are checked in profiles/.
//   imadw_only  : IMAD.WIDE (RZ addend), products folded with one LOP3 per 2 products -> multiplier-pipe rate
//   imad32_acc  : IMAD (32-bit, fused accumulate)                                      -> 32-bit IMAD rate
//   block19     : the block28 inner loop: 19x19 mad.wide.s32 into 37 64-bit columns, operands from smem
//                 (ptxas emits IMAD.WIDE RZ + IADD3/IADD3.X) -> best-case MAC rate of the code shape we ship
//   dfma        : DFMA, for reference
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_imadw_only(unsigned long long* out, unsigned seed, int iters) {
    // products feed back as the next operands (a[k] <- hi, x <- lo of the last one): no ALU work at all
    unsigned a[8], x = seed + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed * (k + 3) + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            unsigned lo[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo[k]), "=r"(a[k]) : "r"(a[k]), "r"(x));
            x = lo[7] | lo[3] | 1u;
        }
    }
    unsigned fold = x;
#pragma unroll
    for (int k = 0; k < 8; k++) fold ^= a[k];
    if (fold == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = fold;
}

__global__ void __launch_bounds__(512) k_imad32_acc(unsigned long long* out, unsigned seed, int iters) {
    unsigned acc[8], x = seed + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = seed * (k + 3) + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(x), "r"(acc[(k + 1) & 7]));
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512) k_dfma(unsigned long long* out, unsigned seed, int iters) {
    double acc[8], x = 1.0 + seed * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k * 0.5 + blockIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[k]) : "d"(x), "d"(acc[(k + 1) & 7]));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    if (s == 0.1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = (unsigned long long)s;
}

__device__ __forceinline__ void madw(long long& acc, int a, int b) { asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b)); }
// the same multiply-accumulate written as a carry pair on the two halves of the accumulator: ptxas turns it into ONE
// IMAD.WIDE Rd, Ra, Rb, Rd (64-bit register addend) instead of IMAD.WIDE .., RZ + IADD3 + IADD3.X
__device__ __forceinline__ void madw_fused(long long& acc, int a, int b) {
    asm("{ .reg .b32 l, h; mov.b64 {l, h}, %0; mad.lo.cc.s32 l, %1, %2, l; madc.hi.s32 h, %1, %2, h; mov.b64 %0, {l, h}; }"
        : "+l"(acc) : "r"(a), "r"(b));
}

// the real inner loop: operands from shared memory ([chunk][lane] int4), 37 column accumulators
template <bool FUSED>
__global__ void __launch_bounds__(256, 2) k_block19(unsigned long long* out, unsigned seed, int iters) {
    extern __shared__ int4 sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * 8 * 5 * 32; i += blockDim.x) sm[i] = make_int4(seed * i, seed + i, i * 7, i ^ seed);
    __syncthreads();
    long long acc[37];
#pragma unroll
    for (int k = 0; k < 37; k++) acc[k] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const int4* ap = sm + ((it & 7) * 5) * 32 + lane;
        const int4* bp = sm + (8 * 5 + ((it * 3) & 7) * 5) * 32 + lane;
        int a[20];
#pragma unroll
        for (int c = 0; c < 5; c++) { int4 v = ap[c * 32]; a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w; }
#pragma unroll
        for (int c = 0; c < 5; c++) {
            int4 bv = bp[c * 32];
            int b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int y = 4 * c + e;
                if (y < 19) {
#pragma unroll
                    for (int x = 0; x < 19; x++) { if (FUSED) madw_fused(acc[x + y], a[x], b4[e]); else madw(acc[x + y], a[x], b4[e]); }
                }
            }
        }
    }
    long long s = 0;
#pragma unroll
    for (int k = 0; k < 37; k++) s ^= acc[k];
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
    unsigned long long* out; CK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * 8));
    CK(cudaFuncSetAttribute(k_block19<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 5 * 32 * 16));
    CK(cudaFuncSetAttribute(k_block19<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 5 * 32 * 16));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    bool first = true;
    auto emit = [&](const char* name, int threads, int cps, double ops, double ms) {
        double per_s = ops / (ms * 1e-3);
        printf("%s  {\"variant\": \"%s\", \"threads_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"op_per_s\": %.6e, \"op_per_clk_per_sm_at_1965MHz\": %.3f}",
               first ? "" : ",\n", name, threads, cps, ms, per_s, per_s / sms / 1.965e9);
        first = false; fflush(stdout);
    };
    for (int warm = 0; warm < 10; warm++) k_imad32_acc<<<sms * 2, 512>>>(out, 12345u, 4000);
    CK(cudaDeviceSynchronize());
    const int iters = 4000;
    for (int threads : {256, 512}) for (int cps : {1, 2, 4}) {
        if (threads * cps > 2048) continue;
        double ms = time_ms([&] { k_imadw_only<<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
        emit("imadw_only", threads, cps, (double)sms * cps * threads * iters * 32, ms);
        ms = time_ms([&] { k_imad32_acc<<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
        emit("imad32_acc", threads, cps, (double)sms * cps * threads * iters * 64, ms);
        ms = time_ms([&] { k_dfma<<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
        emit("dfma", threads, cps, (double)sms * cps * threads * iters * 64, ms);
    }
    for (int cps : {1, 2}) {
        double ms = time_ms([&] { k_block19<false><<<sms * cps, 256, 2 * 8 * 5 * 32 * 16>>>(out, 12345u, 2000); }, 5);
        emit("block19_mac", 256, cps, (double)sms * cps * 256 * 2000 * 361, ms);
        ms = time_ms([&] { k_block19<true><<<sms * cps, 256, 2 * 8 * 5 * 32 * 16>>>(out, 12345u, 2000); }, 5);
        emit("block19_mac_fused", 256, cps, (double)sms * cps * 256 * 2000 * 361, ms);
    }
    {   // sustained: ~2 s of block19 back to back
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        double ops = 0; CK(cudaEventRecord(e0));
        for (int i = 0; i < 60; i++) { k_block19<false><<<sms * 2, 256, 2 * 8 * 5 * 32 * 16>>>(out, 7u, 20000); ops += (double)sms * 2 * 256 * 20000 * 361; }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        emit("block19_mac_sustained", 256, 2, ops, ms);
    }
    printf("\n]}\n");
    return 0;
}
