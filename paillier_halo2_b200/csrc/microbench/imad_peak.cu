// imad_peak.cu — integer-multiply pipe microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the Paillier hot path is bound by the
// 32x32->64 multiply-accumulate rate (SASS IMAD.WIDE.U32[.X]).  This program measures the
// sustained MAC/s of the forms the Barrett kernels actually issue, under load, and prints one JSON
// object; bench.py reads profiles/imad_peak_r01.json (a committed copy of this output) as
// P_imad, the roofline denominator.
//
// Variants (each: grid = 148*k CTAs, long unrolled loops, CUDA-event timed):
//   wide_indep  : mad.wide.u32 acc64 += a*b, ILP independent accumulators          (IMAD.WIDE.U32)
//   wide_chain  : mad.lo.cc/madc.hi.cc rows of 8 pairs + addc (the real inner loop)  (IMAD.WIDE.U32.X)
//   lo32        : mad.lo.u32                                                        (IMAD)
//   chain_alu   : wide_chain + one LOP3/IADD3 per IMAD.WIDE (does the alu pipe co-issue for free?)
//   chain_lds   : wide_chain + 8 LDS.128 per 256 MACs (the block-product operand traffic)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(1024) k_wide_indep(uint64_t* out, uint32_t seed, int iters) {
    uint64_t acc[ILP];
    uint32_t x = seed + threadIdx.x, y = seed * 3 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = k + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < ILP; k++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    if (s == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024) k_lo32(uint64_t* out, uint32_t seed, int iters) {
    uint32_t acc[ILP];
    uint32_t x = seed + threadIdx.x, y = seed * 3 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = k + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < ILP; k++)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(x), "r"(y));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= acc[k];
    if (s == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one carry-chained row: acc[0..2P] += a[0,2,..,2P-2] * b, carry absorbed in acc[2P]
template <int P>
__device__ __forceinline__ void row(uint32_t* acc, const uint32_t* a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                 : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
    for (int i = 1; i < P; i++)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(acc[2 * i]), "+r"(acc[2 * i + 1]) : "r"(a[2 * i]), "r"(b));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[2 * P]));
}

// MODE 0: pure chain; 1: + one alu op per IMAD.WIDE; 2: + 8 LDS.128 per 256 MAC
template <int MODE>
__global__ void __launch_bounds__(512) k_wide_chain(uint64_t* out, uint32_t seed, int iters) {
    __shared__ uint4 sm[512 * 2];
    uint32_t a[16], b[16], E[34], O[34];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed * (i + 7) + blockIdx.x; }
#pragma unroll
    for (int i = 0; i < 34; i++) { E[i] = i; O[i] = i + seed; }
    sm[threadIdx.x] = make_uint4(a[0], a[1], a[2], a[3]);
    sm[threadIdx.x + 512] = make_uint4(b[0], b[1], b[2], b[3]);
    __syncthreads();
    uint32_t junk = seed;
    for (int it = 0; it < iters; it++) {
        if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 v = sm[(threadIdx.x + q * 32 + it) & 511];
                uint4 w = sm[512 + ((threadIdx.x + q * 32 + it) & 511)];
                a[4 * q] ^= v.x; a[4 * q + 1] ^= v.y; a[4 * q + 2] ^= v.z; a[4 * q + 3] ^= v.w;
                b[4 * q] ^= w.x; b[4 * q + 1] ^= w.y; b[4 * q + 2] ^= w.z; b[4 * q + 3] ^= w.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            row<8>(E + j, a, b[j]);
            row<8>(O + j, a + 1, b[j]);
            row<8>(O + j, a, b[j + 1]);
            row<8>(E + j + 2, a + 1, b[j + 1]);
            if (MODE == 1) {
#pragma unroll
                for (int q = 0; q < 32; q++)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(junk) : "r"(a[q & 15]), "r"(b[q & 15]));
            }
        }
    }
    uint32_t s = junk;
#pragma unroll
    for (int i = 0; i < 34; i++) s ^= E[i] ^ O[i];
    if (s == 0x1234567) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Result { const char* name; int threads; int ctas_per_sm; double gmacs; double ms; };

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    uint64_t* out; CK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * 8));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_rate_khz_attr\": %d, \"results\": [\n", p.name, sms, clk_khz);
    bool first = true;
    auto emit = [&](const char* name, int threads, int cps, double macs, double ms) {
        double per_s = macs / (ms * 1e-3);
        printf("%s  {\"variant\": \"%s\", \"threads_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"mac_per_s\": %.6e, \"mac_per_clk_per_sm_at_1965MHz\": %.3f}",
               first ? "" : ",\n", name, threads, cps, ms, per_s, per_s / sms / 1.965e9);
        first = false; fflush(stdout);
    };
    const int iters = 2000;
    // sustained: a long run (~seconds) first to get to steady clocks
    for (int warm = 0; warm < 20; warm++) k_wide_indep<8><<<sms * 2, 512>>>(out, 12345u, iters);
    CK(cudaDeviceSynchronize());
    for (int threads : {128, 256, 512, 1024}) {
        for (int cps : {1, 2}) {
            if (threads * cps > 2048) continue;
            double ms = time_ms([&] { k_wide_indep<8><<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
            emit("wide_indep_ilp8", threads, cps, (double)sms * cps * threads * iters * 8 * 8, ms);
            ms = time_ms([&] { k_wide_indep<2><<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
            emit("wide_indep_ilp2", threads, cps, (double)sms * cps * threads * iters * 8 * 2, ms);
            ms = time_ms([&] { k_wide_indep<1><<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
            emit("wide_indep_ilp1", threads, cps, (double)sms * cps * threads * iters * 8 * 1, ms);
            ms = time_ms([&] { k_lo32<8><<<sms * cps, threads>>>(out, 12345u, iters); }, 5);
            emit("lo32_ilp8", threads, cps, (double)sms * cps * threads * iters * 8 * 8, ms);
        }
    }
    const int citers = 400;
    for (int threads : {128, 256, 512}) {
        for (int cps : {1, 2, 4}) {
            if (threads * cps > 1024) continue;
            double macs = (double)sms * cps * threads * citers * 256;
            double ms = time_ms([&] { k_wide_chain<0><<<sms * cps, threads>>>(out, 12345u, citers); }, 5);
            emit("wide_chain", threads, cps, macs, ms);
            ms = time_ms([&] { k_wide_chain<1><<<sms * cps, threads>>>(out, 12345u, citers); }, 5);
            emit("wide_chain_alu1to1", threads, cps, macs, ms);
            ms = time_ms([&] { k_wide_chain<2><<<sms * cps, threads>>>(out, 12345u, citers); }, 5);
            emit("wide_chain_lds", threads, cps, macs, ms);
        }
    }
    // sustained figure: 3 s of the chained kernel back to back (power/clock steady state)
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int launches = 0; double macs = 0;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 300; i++) { k_wide_chain<0><<<sms * 2, 512>>>(out, 777u, 2000); launches++; macs += (double)sms * 2 * 512 * 2000 * 256; }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        emit("wide_chain_sustained", 512, 2, macs, ms);
    }
    printf("\n]}\n");
    return 0;
}
