// block28u.cuh — "block28u": the two constant-operand phases of the block28 modular multiplication on the 5th-generation tensor
// core (tcgen05.mma kind::i8, accumulators in TMEM), phase A unchanged on the IMAD pipe.
//
// Same arithmetic as block28t (block28.cuh, phase_mma / qhat_to_bytes / low_to_value) — the lazy digits it leaves in V are
// bit-identical, tests/model_block28.py is the model of both — but the GEMM  C[lane][p] = sum_k A7[lane][k] K7[p - k]  is issued by
// ONE thread and runs asynchronously on the tensor core.
//
// How 32 or 64 ciphertexts fill a 128-row MMA.  The per-lane s8 rows are stored K-major without swizzle as
// [16-byte K chunk][row][16], ROWS = 32 LG rows per chunk (LG = lane groups of 32 ciphertexts per CTA: 1 or 2).  A UMMA descriptor
// with 128 B between 8-row groups then makes MMA rows ROWS j .. ROWS j + ROWS - 1 the SAME rows read j chunks further along K.  By
// the Toeplitz structure a row shifted by 16 j in k holds the output columns shifted by 16 j in p:
// D[ROWS j + row][n] = C[row][p_hi - n + 16 j].  One MMA therefore yields TN + 16 (NSH - 1) distinct output columns per ciphertext
// (NSH = 128 / ROWS copies: 4 or 2) and every TMEM lane quadrant holds ciphertexts, so all warps of the CTA take part in the fold
// (a warp reads only the quadrant warp % 4).  Zero chunks in front of and behind the rows stand in for k < 0 and k >= K7.
// LG = 2 halves the tensor-core time and the shared-memory traffic per ciphertext (the B operand and the instruction are shared
// by 64 ciphertexts): measured, the MMAs run at the speed the tensor core can fetch its operands from shared memory (~90 B/clk).
//
// The Toeplitz operand is never materialised: with the tile's columns taken in DECREASING p, core matrix (column group g, K chunk
// j) of any (tile, k-step) is entry u0 + g + 2j of one per-key table CM[u][r][b] = Rev[8u + r + b] (descriptor strides 128 B along
// N, 256 B along K) — csrc/microbench/umma_toeplitz.cu measured and verified this form.
//
// Tiles: G ranges per ciphertext, one per warp of its lane group — 8 ranges of 5 digits (160 columns) with 8 warps, 16 ranges of 3
// digits (192 columns) with 16; a warp folds one range from its TMEM quadrant and reads the 4 columns below its range itself for
// the incoming carry, so tiles and ranges are independent of each other.
// Two TMEM buffers per CTA; a warp that has read its columns out arrives on the buffer's "empty" mbarrier, one thread of warp 0 waits for the
// arrivals and issues the MMAs of the tile after next — no CTA barrier between tiles.
//
// Shared memory, LG = 1 (|n| = 2048: 111 KB, two CTAs per SM; 3072 / 4096: 166 / 211 KB, one): V | B | T(2L) | tail of the q-hat
// rows | constants | CM(mu) | CM(Nt).
// LG = 2 (190 KB, one CTA of 16 warps per SM): V0 B0 V1 B1 | T0lo T1lo T0hi T1hi | tail | constants | CM | CM.  The q1 rows overlay
// V, B (dead after phase A), the q-hat rows overlay the upper halves of T (dead once q1 has been cut out); phase A's stash of Hi
// digits, which block28t keeps in the Q buffer, lives in TMEM columns (tcgen05.st / ld); the low-part sums of phase C go to a flat
// array in B.  Everything the phases B / C execute on the CUDA cores avoids the FMA pipes (see "Opaque operands").
#pragma once
#include "block28.cuh"
#include <type_traits>

namespace pb200 {
namespace b28 {

// WIT: layout of the witness engine — the q1 rows overlay B and a gap behind it instead of V | B, because the witness step must
// preserve its first operand (block28.cuh, mulmod_w)
template <class C, int LG, bool WIT = false>
struct UL {
    static constexpr int L = C::L, K7 = C::K7, G = C::G;
    static constexpr int KCH = K7 / 16;                        // 16-byte K chunks per row
    static constexpr int ROWS = 32 * LG, CHB = 16 * ROWS;      // rows and bytes per chunk
    static constexpr int NSH = 128 / ROWS;                     // shifted copies of the rows in one MMA
    static constexpr int SHIFTC = 16 * (NSH - 1);              // extra output columns the copies bring
    static constexpr int FRONT = 2 * ((SHIFTC + 31) / 32), BACK = NSH - 1;      // zero chunks before / after the row
    static constexpr int KOFF = 16 * FRONT;
    static constexpr int ACH = FRONT + KCH + BACK;
    static constexpr int A_BYTES = ACH * CHB;
    // a tile is G ranges per ciphertext (one per warp of the lane group); a range is RD digits: 5 with 8 warps, 3 with 16
    static constexpr int RD = G == 16 ? 3 : 5, RCOLS = 4 * RD, TCOLS = G * RCOLS, DT = TCOLS / 4;
    static constexpr int TN = ((TCOLS - SHIFTC + 4) + 15) / 16 * 16;      // MMA N: the columns the unshifted rows must see
    static constexpr int NWARPS = C::G * LG, THREADS = 32 * NWARPS;
    static constexpr int RPS = G / NSH;                        // ranges per (ciphertext group, shift) = warps per TMEM quadrant
    static constexpr int P_BASE_H = 4 * (L - 2);               // phase B keeps two guard digits below q-hat
    static constexpr int NT_H = (4 * (L + 2) + TCOLS - 1) / TCOLS, NT_L = (4 * L + TCOLS - 1) / TCOLS;
    static constexpr int TMEM_COLS = 2 * TN <= 256 ? 256 : 512;
    // mbarrier completions of buffer b per phase: "full" once per tile, "empty" once per tile that is followed by another tile in
    // the same buffer; over one multiplication each must be even so that the wait parities are compile-time constants
    __host__ __device__ static constexpr int n_full(bool high, int b) { return ((high ? NT_H : NT_L) + 1 - b) / 2; }
    __host__ __device__ static constexpr int n_empty(bool high, int b) { return ((high ? NT_H : NT_L) - 2 + 1 - b) / 2 > 0 ? ((high ? NT_H : NT_L) - 2 + 1 - b) / 2 : 0; }
    __host__ __device__ static constexpr unsigned par_full(bool high, int s) { return (unsigned)(((high ? 0 : n_full(true, s & 1)) + (s >> 1)) & 1); }
    __host__ __device__ static constexpr unsigned par_empty(bool high, int s) { return (unsigned)(((high ? 0 : n_empty(true, s & 1)) + (s >> 1)) & 1); }
    __host__ __device__ static constexpr int p_top(bool high, int t) { return (high ? P_BASE_H + NT_H * TCOLS : NT_L * TCOLS) - 1 - TCOLS * t; }
    __host__ __device__ static constexpr int p_hi(bool high, int t) { return p_top(high, t) - SHIFTC; }
    // k range of a tile in the coordinates of the unshifted rows: the band of its TN columns, extended below zero for the shifted rows
    __host__ __device__ static constexpr int k_start(int ph) {
        int k_lo = ph - (TN - 1) - (K7 - 1);
        if (k_lo < -SHIFTC) k_lo = -SHIFTC;
        return ((k_lo + KOFF) / 32) * 32 - KOFF;
    }
    __host__ __device__ static constexpr int n_ksteps(int ph) { return ((ph < K7 - 1 ? ph : K7 - 1) - k_start(ph)) / 32 + 1; }
    // Rev[z] = K7c[z0 - z]; z0 = 7 (mod 8) like every p_hi, and >= the largest p_hi - k any tile touches
    __host__ __device__ static constexpr int z0(bool high) {
        int m = 0;
        for (int t = 0; t < (high ? NT_H : NT_L); t++) { const int d = p_hi(high, t) - k_start(p_hi(high, t)); if (d > m) m = d; }
        return m + ((7 - m % 8) + 8) % 8;
    }
    __host__ __device__ static constexpr int ncm(bool high) {
        int m = 0;
        for (int t = 0; t < (high ? NT_H : NT_L); t++) {
            const int ph = p_hi(high, t), k_last = k_start(ph) + 32 * (n_ksteps(ph) - 1);
            const int u = (z0(high) - ph + k_last) / 8 + TN / 8 + 2;
            if (u > m) m = u;
        }
        return m;
    }
    static constexpr int Z0_H = z0(true), Z0_L = z0(false), NCM_H = ncm(true), NCM_L = ncm(false);
    // byte offsets in dynamic shared memory
    static constexpr int VALB = C::VAL4 * 16;
    static constexpr int ASB_OFF = WIT ? VALB : 0;                                  // where the q1 rows start
    // the gap also holds the witness tail's per-block hash sums (G * 32 u64)
    static constexpr int GAP = WIT ? ((A_BYTES - VALB > G * 32 * 8 ? A_BYTES - VALB : G * 32 * 8) + 127) / 128 * 128 : 0;
    static constexpr int OFF_T = 2 * LG * VALB + GAP, OFF_ASC = OFF_T + LG * VALB;   // T low halves of all groups, then the high halves
    static constexpr int T_HI_JUMP = (LG - 1) * VALB;                                // block d >= G of a group lies this much further
    static constexpr int END_ASC = OFF_ASC + A_BYTES;
    static constexpr int OFF_CONST = ((END_ASC > OFF_T + 2 * LG * VALB ? END_ASC : OFF_T + 2 * LG * VALB) + 127) / 128 * 128;
    static constexpr int CONST_BYTES = (3 * C::ENTRY4 * 16 + 127) / 128 * 128;
    static constexpr int OFF_CMH = OFF_CONST + CONST_BYTES, OFF_CML = OFF_CMH + NCM_H * 128, OFF_BAR = OFF_CML + NCM_L * 128;
    static constexpr int KEY_BYTES = OFF_BAR - OFF_CONST;      // per-key image copied from global memory: constants, CM(mu), CM(Nt)
    static constexpr int OFF_EST = OFF_BAR + 64;               // witness engine: one int per lane (k estimate)
    static constexpr size_t SMEM_BYTES = (size_t)OFF_BAR + 64 + (WIT ? 128 : 0);
    static constexpr int CTAS_PER_SM = (int)((233472 / (SMEM_BYTES + 1024)) < (512 / THREADS) ? (233472 / (SMEM_BYTES + 1024)) : (512 / THREADS));
    // compiled for configurations with whole k-steps per row, tile counts for which every mbarrier completes an even number of times
    // per multiplication (the wait parities are compile-time constants), the q1 rows inside V | B, and 8 or 16 warps per lane group
    static constexpr bool SUPPORTED = (K7 % 32 == 0) && (NT_H >= 2) && (NT_L >= 2) &&
                                      ((n_full(true, 0) + n_full(false, 0)) % 2 == 0) && ((n_full(true, 1) + n_full(false, 1)) % 2 == 0) &&
                                      ((n_empty(true, 0) + n_empty(false, 0)) % 2 == 0) && ((n_empty(true, 1) + n_empty(false, 1)) % 2 == 0) &&
                                      (WIT ? (LG == 1 && A_BYTES <= VALB + GAP && GAP >= G * 32 * 8) : (A_BYTES <= 2 * LG * VALB)) &&
                                      ((G == 8) || (G == 16 && LG == 1)) && (TCOLS % 16 == 0) && (RCOLS + 4 <= 24) &&
                                      (Z0_H % 8 == 7) && (Z0_L % 8 == 7) && (P_BASE_H % 8 == 0) && (TN <= 256) &&
                                      (2 * TN <= TMEM_COLS) && CTAS_PER_SM >= 1 && SMEM_BYTES <= 232448;
};

// Shared-memory view of one thread's lane group.  Member names follow Smem<C> so the kernels are written once.
template <class C, int LG, bool WIT = false>
struct SmemU {
    int4* V; int4* B; int4* T; const int4* mu; const int4* Nt; const int4* two_sh;
    unsigned char* base;
    uint32_t tmem;          // TMEM base address of this CTA's columns
    int group;              // lane group of this thread in phase A and in the kernels' value handling: warp / G
    __device__ __forceinline__ SmemU(int4* b) {
        typedef UL<C, LG, WIT> U;
        base = (unsigned char*)b;
        group = LG == 1 ? 0 : (int)(threadIdx.x >> 5) / C::G;
        V = b + group * 2 * C::VAL4; B = V + C::VAL4;
        T = (int4*)(base + U::OFF_T) + group * C::VAL4;
        const int4* k = (const int4*)(base + U::OFF_CONST);
        mu = k; Nt = k + C::ENTRY4; two_sh = k + 2 * C::ENTRY4;
        tmem = 0;
    }
    // block d of this group's 2L-digit product
    __device__ __forceinline__ int4* tblk(int d, int lane) const { return T + d * C::BLK4 + (d >= C::G ? UL<C, LG, WIT>::T_HI_JUMP / 16 : 0) + lane; }
    __device__ __forceinline__ unsigned char* asb() const { return base + UL<C, LG, WIT>::ASB_OFF; }
    __device__ __forceinline__ unsigned char* asc() const { return base + UL<C, LG, WIT>::OFF_ASC; }
    __device__ __forceinline__ uint64_t* bars() const { return (uint64_t*)(base + UL<C, LG, WIT>::OFF_BAR); }
    __device__ __forceinline__ volatile uint32_t* slots() const { return (volatile uint32_t*)(base + UL<C, LG, WIT>::OFF_BAR + 32); }   // [0] TMEM base, [1] dead
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, K-major, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
// one lane of a fully active warp (elect.sync): code under it is single-threaded to ptxas, which then issues tcgen05.mma without the
// per-thread serialisation loop it wraps around a warp-level instruction in merely divergent code
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Bounded wait: a wrong descriptor must not hang the box.  The first time-out marks the CTA dead (every later wait returns at
// once) and the kernel traps when it ends, so the failure is loud and the launch still terminates.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* dead) {
    const uint32_t a = smem_u32(bar);
    for (int spin = 0; spin < (1 << 21); spin++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (*dead) return;
    }
    *dead = 1;
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const int* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const int* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// barrier of one lane group (8 warps); with one group per CTA it is the CTA barrier
template <int LG, int NTHREADS>
__device__ __forceinline__ void group_sync(int group) {
    if (LG == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(NTHREADS) : "memory");
}

// ---- per-CTA set-up / tear-down -----------------------------------------------------------------------------------------
template <class C, int LG, bool WIT>
__device__ __forceinline__ void umma_setup(SmemU<C, LG, WIT>& S) {
    using U = UL<C, LG, WIT>;
    const int warp = threadIdx.x >> 5;
    // bars 0, 1: the MMAs of a TMEM buffer are complete (tcgen05.commit);  2, 3: every warp has read the buffer out (one arrival per warp)
    if (threadIdx.x == 0) {
        mbar_init(&S.bars()[0], 1); mbar_init(&S.bars()[1], 1); mbar_init(&S.bars()[2], U::NWARPS); mbar_init(&S.bars()[3], U::NWARPS);
        S.slots()[1] = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32((const void*)S.slots())), "n"(U::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // the zero chunks behind the q-hat rows lie beyond T and are never written again
    for (int i = threadIdx.x; i < U::BACK * U::ROWS; i += U::THREADS)
        *(int4*)(S.asc() + (U::FRONT + U::KCH) * U::CHB + i * 16) = make_int4(0, 0, 0, 0);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    S.tmem = S.slots()[0];
}
template <class C, int LG, bool WIT>
__device__ __forceinline__ void umma_teardown(SmemU<C, LG, WIT>& S) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(S.tmem), "n"(UL<C, LG, WIT>::TMEM_COLS) : "memory");
    }
    if (S.slots()[1]) {
        if (threadIdx.x == 0) printf("pb200 block28u: tcgen05 completion never arrived (CTA %d)\n", (int)blockIdx.x);
        __trap();
    }
}

// ---- phase A with the stash in TMEM --------------------------------------------------------------------------------------
// phase_product of block28.cuh for this thread's lane group; the Hi digits of a warp's first anti-diagonal wait for the merge in
// TMEM columns [32 (warp / 4), + CH * 4) of the warp's own lane quadrant instead of the Q buffer.
template <class C, int LG, bool WIT = false>
__device__ __noinline__ void phase_product_u(int4* smem_base, const int4* Y, int sqr, uint32_t tmem) {
    constexpr int G = C::G, BL = C::BL;
    SmemU<C, LG, WIT> S(smem_base);
    const int lane = threadIdx.x & 31, cwarp = threadIdx.x >> 5, warp = cwarp % G;      // warp: role within the lane group
    const unsigned v_addr = (unsigned)__cvta_generic_to_shared(S.V) + lane * 16;
    const unsigned y_addr = sqr ? v_addr : (unsigned)__cvta_generic_to_shared(Y) + lane * 16;
    const uint32_t stash = tmem + ((uint32_t)((cwarp & 3) * 32) << 16) + (uint32_t)((cwarp >> 2) * 32);
    Pending<C> pend;
    pend.blk = -1;
    int blk0 = -1, spill0 = 0;
    const int njobs = warp < G - 1 ? 2 : 1;
#pragma unroll 1
    for (int half = 0; half < njobs; half++) {
        const int d = warp + half * G;
        const int i_lo = half ? d - G + 1 : 0;
        const int i_hi = sqr ? d / 2 : (half ? G - 1 : d);
        long long carry;
        {   // columns 0 .. BL-1 -> Lo digits
            long long acc[BL];
#pragma unroll
            for (int k = 0; k < BL; k++) acc[k] = 0;
            job_half<C, false>(acc, d, i_lo, i_hi, sqr, v_addr, y_addr);
            int lo[C::CH * 4];
            carry = ripple_cols<C>(acc, BL, 0, lo);
#pragma unroll
            for (int k = BL; k < C::CH * 4; k++) lo[k] = 0;
            store_block<C>(S.tblk(d, lane), lo);
        }
        {   // columns BL .. 2BL-2 -> Hi digits, last digit and spill from the final carry
            long long acc[BL];
#pragma unroll
            for (int k = 0; k < BL; k++) acc[k] = 0;
            job_half<C, true>(acc, d, i_lo, i_hi, sqr, v_addr, y_addr);
            carry = ripple_cols<C>(acc, BL - 1, carry, pend.hi);
            int dtop = sgxt28((int)carry);
            pend.hi[BL - 1] = dtop;
            pend.spill = (int)((carry - dtop) >> W);
#pragma unroll
            for (int k = BL; k < C::CH * 4; k++) pend.hi[k] = 0;
        }
        pend.carry = 0;
        pend.blk = d;
        if (half == 0 && njobs == 2) {          // warp-uniform
            tmem_st16(stash, pend.hi);
            if (C::CH * 4 > 16) tmem_st4(stash + 16, pend.hi + 16);
            tmem_wait_st();
            blk0 = d; spill0 = pend.spill;
        }
    }
    if (warp == G - 1) {                        // block 2G-1 has no Lo contribution
        int z[C::CH * 4];
#pragma unroll
        for (int k = 0; k < C::CH * 4; k++) z[k] = 0;
        store_block<C>(S.tblk(2 * G - 1, lane), z);
    }
    group_sync<LG, C::THREADS>(S.group);          // the lane groups of a CTA run phase A independently of each other
    int carry0 = 0;
    if (blk0 >= 0) {
        int h0[C::CH * 4];
        tmem_ld16(stash, h0);
        if (C::CH * 4 > 16) tmem_ld4(stash + 16, h0 + 16);
        tmem_wait_ld();
        int4* p = S.tblk(blk0 + 1, lane);
        carry0 = add_ripple_block<C>(p, p, h0, 1);
    }
    {
        int4* p = S.tblk(pend.blk + 1, lane);
        pend.carry = add_ripple_block<C>(p, p, pend.hi, 1);
    }
    group_sync<LG, C::THREADS>(S.group);
    if (blk0 >= 0) *(int*)S.tblk(blk0 + 2, lane) += carry0 + spill0;
    if (pend.blk + 2 <= 2 * G - 1) *(int*)S.tblk(pend.blk + 2, lane) += pend.carry + pend.spill;
    __syncthreads();          // CTA-wide: the q1 rows that phases_bc_umma writes next overlay V and B of every lane group
}

// ---- phases B and C ---------------------------------------------------------------------------------------------------------
// BL packed words of a row, starting at word w0 of the row, into the chunked layout: 128-bit stores wherever a chunk is whole
template <int BL, int LEAD, int CHB>
__device__ __forceinline__ void store_row_u_aligned(unsigned char* rowbase, int w0, const unsigned (&w)[BL]) {
#pragma unroll
    for (int k = 0; k < LEAD; k++) *(unsigned*)(rowbase + ((w0 + k) >> 2) * CHB + ((w0 + k) & 3) * 4) = w[k];
    constexpr int NV = (BL - LEAD) / 4;
    const int c0 = (w0 + LEAD) >> 2;
#pragma unroll
    for (int c = 0; c < NV; c++)
        *(uint4*)(rowbase + (c0 + c) * CHB) = make_uint4(w[LEAD + 4 * c], w[LEAD + 4 * c + 1], w[LEAD + 4 * c + 2], w[LEAD + 4 * c + 3]);
#pragma unroll
    for (int k = LEAD + 4 * NV; k < BL; k++) *(unsigned*)(rowbase + ((w0 + k) >> 2) * CHB + ((w0 + k) & 3) * 4) = w[k];
}
template <int BL, int CHB>
__device__ __forceinline__ void store_row_u(unsigned char* rowbase, int w0, const unsigned (&w)[BL]) {
    switch ((4 - (w0 & 3)) & 3) {          // warp-uniform
        case 0: store_row_u_aligned<BL, 0, CHB>(rowbase, w0, w); break;
        case 1: store_row_u_aligned<BL, 1, CHB>(rowbase, w0, w); break;
        case 2: store_row_u_aligned<BL, 2, CHB>(rowbase, w0, w); break;
        default: store_row_u_aligned<BL, 3, CHB>(rowbase, w0, w); break;
    }
}

// Opaque operands.  Phases B and C share the SM with the other CTA's phase A, which saturates the FMA pipes with IMAD.WIDE, and ptxas
// likes to turn constant left shifts, two-input adds and moves into IMAD.SHL / IMAD.IADD / IMAD.MOV — every one of them then queues
// behind eight warps of IMAD.WIDE (measured: phases B + C 16.7 k clk alone, 29 k next to a phase-A CTA).  A shift count or a zero
// addend read from constant memory cannot be folded away, so these become SHF and IADD3 on the ALU pipe.
static __constant__ int c_opq[8] = {0, 1, 2, 3, 7, 9, 14, 28};
#define OPQ_Z (c_opq[0])
#define OPQ_1 (c_opq[1])
#define OPQ_2 (c_opq[2])
#define OPQ_3 (c_opq[3])
#define OPQ_7 (c_opq[4])
#define OPQ_14 (c_opq[6])
#define OPQ_28 (c_opq[7])
__device__ __forceinline__ unsigned lop3_sel(unsigned a, unsigned b, unsigned m) {          // (a & ~m) | (b & m)
    unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(m), "r"(b), "r"(a)); return d;      // m ? b : a
}
__device__ __forceinline__ unsigned lop3_xor_or(unsigned a, unsigned b, unsigned c) {       // (a ^ b) | c
    unsigned d; asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
// split7_pack of block28.cuh on dp = d + (64 | 64 << 7 | 64 << 14), shifts and adds on the ALU pipe
__device__ __forceinline__ unsigned split7_pack_biased(unsigned dp) {
    const unsigned s1 = dp << OPQ_1, s2 = dp << OPQ_2, s3 = dp << OPQ_3;
    unsigned x = lop3_sel(dp, s1, 0x7F00u);
    x = lop3_sel(x, s2, 0x7F0000u) & 0x7F7F7Fu;
    const unsigned y = x + 0x404040u + (unsigned)OPQ_Z;
    return lop3_xor_or(y, 0x808080u, s3 & 0xFF000000u);
}
constexpr unsigned SPLIT_BIAS = 64u | (64u << 7) | (64u << 14);

// v[0..3] = columns c3, c2, c1, c0 of one digit:  c0 + c1 2^7 + c2 2^14 + c3 2^21 = (lob - 2^27) + ca 2^28,  lob in [0, 2^28)
__device__ __forceinline__ void fold4(const int* v, int& lob, int& ca) {
    const int z = OPQ_Z;
    const int t2 = v[3] + (v[2] << OPQ_7) + (1 << (W - 1));
    const int highp = v[1] + (v[0] << OPQ_7) + z;
    const int t5 = (highp << OPQ_14) & 0x0FFFC000;
    const int tb = t2 + t5 + z;
    lob = tb & ((1 << W) - 1);
    ca = (tb >> W) + (highp >> 14) + z;                                  // carry into the digit above
}

// all MMAs of issue slot s (HIGH: tiles from the bottom up, LOW: from the top down — longest k range first) into TMEM buffer s & 1
template <class C, int LG, bool WIT, bool HIGH>
__device__ __forceinline__ void umma_issue(const SmemU<C, LG, WIT>& S, int s) {
    using U = UL<C, LG, WIT>;
    const int t = HIGH ? U::NT_H - 1 - s : s;
    const int ph = U::p_hi(HIGH, t), ks0 = U::k_start(ph), nks = U::n_ksteps(ph);
    const uint32_t a_base = smem_u32(HIGH ? S.asb() : S.asc()), cm_base = smem_u32(S.base + (HIGH ? U::OFF_CMH : U::OFF_CML));
    const uint32_t d_tmem = S.tmem + (uint32_t)((s & 1) * U::TN);
    constexpr int Z0 = HIGH ? U::Z0_H : U::Z0_L;
    constexpr uint32_t idesc = umma_idesc(128, U::TN);
    // descriptors of the first k-step; a k-step further is two K chunks of A and four table entries of B (512 B)
    uint64_t a_desc = umma_desc(a_base + (uint32_t)((ks0 + U::KOFF) >> 4) * (uint32_t)U::CHB, U::CHB, 128);
    uint64_t b_desc = umma_desc(cm_base + (uint32_t)((Z0 - ph + ks0) >> 3) * 128u, 256, 128);
    umma_i8(d_tmem, a_desc, b_desc, 0u, idesc);
#pragma unroll 4
    for (int ks = 1; ks < nks; ks++) {
        a_desc += (2 * U::CHB) >> 4; b_desc += 512 >> 4;
        umma_i8(d_tmem, a_desc, b_desc, 1u, idesc);
    }
    umma_commit(&S.bars()[s & 1]);
}

// fold of one range of a tile: 5 digits (+ the digit below for its carry).  HIGH: packed q-hat words into the rows of phase C;
// LOW: lo(q-hat Nt) digit sums into the flat array F[digit][lane] (the group's B buffer), subtracted from T by the ripple pass.
// dst: the address of the range's top digit (HIGH: word 0 of its 16-byte chunk in this row; LOW: its F entry)
template <class C, int CHB, int RD, bool HIGH>
__device__ __forceinline__ void umma_fold(const int* v /* 4 (RD + 1) columns, top digit first */, unsigned char* dst, int d_top,
                                          int r /* HIGH: (d_top - 2) & 3, warp-uniform */) {
    constexpr int L = C::L;
    int lob[RD + 1], ca[RD + 1];
#pragma unroll
    for (int e = 0; e <= RD; e++) fold4(v + 4 * e, lob[e], ca[e]);
    if (HIGH) {
        // q-hat digit jj - 2 = LO[jj] + CA[jj], not rippled (block28.cuh, qhat_to_bytes); the two guard digits jj = 0, 1 (the last two
        // digits of the phase's bottom range, d_top == RD - 1) only feed a carry into q-hat digit 0
        int carry_g = 0;
        if (d_top == RD - 1) { const int t1 = lob[RD - 2] + ca[RD - 1] - (1 << (W - 1)); carry_g = (t1 - sgxt28(t1)) >> W; }
        unsigned w[RD];
#pragma unroll
        for (int e = 0; e < RD; e++)
            w[e] = split7_pack_biased((unsigned)(lob[e] + ca[e + 1] + (int)(SPLIT_BIAS - (1u << (W - 1))) + (e == RD - 3 ? carry_g : 0)));
        const int qd0 = d_top - 2;
        auto put = [&](auto RC) {
            constexpr int R = decltype(RC)::value;
#pragma unroll
            for (int e = 0; e < RD; e++) {
                const int wd = R - e, ch = wd >= 0 ? wd / 4 : -((3 - wd) / 4), word = wd - 4 * ch;
                if ((unsigned)(qd0 - e) < (unsigned)L) *(unsigned*)(dst + ch * CHB + word * 4) = w[e];
            }
        };
        switch (r) {          // warp-uniform: every store gets an immediate offset
            case 0: put(std::integral_constant<int, 0>{}); break;
            case 1: put(std::integral_constant<int, 1>{}); break;
            case 2: put(std::integral_constant<int, 2>{}); break;
            default: put(std::integral_constant<int, 3>{}); break;
        }
    } else {
#pragma unroll
        for (int e = 0; e < RD; e++)
            if (d_top - e < L) *(int*)(dst - e * 128) = lob[e] + ca[e + 1] - (1 << (W - 1));
    }
}

// this warp's 4 (RD + 1) accumulator columns of a TMEM buffer
template <int RD>
__device__ __forceinline__ void umma_load(uint32_t ta, int* v) {
    tmem_ld16(ta, v);
    if (RD == 5) tmem_ld8(ta + 16, v + 16);
    tmem_wait_ld();
}

// q1 = T digits [L-1, 2L-1) as s8 rows; Q = hi(q1 mu); V = ripple(lo(T) - lo(Q Nt)).
// WIT: stops after the low-part sums are in F (the exact tail of the witness step, w_tail_u, takes over from there)
// EL: the MMAs are issued by the lane elect.sync picks (3 instructions per MMA) instead of by thread 0 in divergent code (9, with
// ptxas' per-thread serialisation loop).  Faster everywhere except in k_tally (5.27 against 5.09 ms), which keeps the old form.
template <class C, int LG, bool WIT = false, bool EL = true>
__device__ __noinline__ void phases_bc_umma(int4* smem_base, uint32_t tmem) {
    using U = UL<C, LG, WIT>;
    constexpr int G = C::G, BL = C::BL;
    SmemU<C, LG, WIT> S(smem_base);
    S.tmem = tmem;
    const int lane = threadIdx.x & 31, cwarp = threadIdx.x >> 5, warp = cwarp % G;
    volatile uint32_t* dead = S.slots() + 1;
    {
        // block `warp` of q1: digit 0 from the top of T block G + warp - 1, the rest from block G + warp (load_q1_block of block28.cuh)
        int t[C::CH * 4], a[BL];
        load_block<C>(t, S.tblk(G + warp, lane), 32);
        const int* below = (const int*)S.tblk(G + warp - 1, lane);
        a[0] = below[((BL - 1) / 4) * 32 * 4 + ((BL - 1) % 4)];
#pragma unroll
        for (int k = 1; k < BL; k++) a[k] = t[k - 1];
        if (warp == G - 1) a[BL - 1] += t[BL - 1] << OPQ_28;       // fold digit 2L-1 (|.| <= 1) into digit 2L-2
        unsigned w[BL];
#pragma unroll
        for (int k = 0; k < BL; k++) w[k] = split7_pack_biased((unsigned)(a[k] + (int)SPLIT_BIAS + OPQ_Z));
        store_row_u<BL, U::CHB>(S.asb() + U::FRONT * U::CHB + (S.group * 32 + lane) * 16, warp * BL, w);
        for (int i = threadIdx.x; i < (U::FRONT + U::BACK) * U::ROWS; i += U::THREADS) {
            const int ch = i / U::ROWS;
            *(int4*)(S.asb() + (ch < U::FRONT ? ch : U::KCH + ch) * U::CHB + (i % U::ROWS) * 16) = make_int4(0, 0, 0, 0);
        }
    }
    fence_async_smem();
    __syncthreads();
    if (cwarp == 0) {
        if (EL ? elect_one() : lane == 0) { tc_fence_after(); umma_issue<C, LG, WIT, true>(S, 0); umma_issue<C, LG, WIT, true>(S, 1); }
        __syncwarp();
    }
    // T's upper halves are dead now: zero chunks in front of the q-hat rows
    for (int i = threadIdx.x; i < U::FRONT * U::ROWS; i += U::THREADS) *(int4*)(S.asc() + i * 16) = make_int4(0, 0, 0, 0);
    // per-thread addressing of the folds.  TMEM quadrant q = warp % 4 holds MMA rows 32 q ..: ciphertext group ge, shift j; the
    // warps of a quadrant share the ranges of (ge, j): range ri of a tile (0 = top)
    const int q = cwarp & 3, ge = LG == 1 ? 0 : (q & 1), j = LG == 1 ? q : (q >> 1);
    const int ri = (U::NSH - 1 - j) * U::RPS + (cwarp >> 2);
    const uint32_t ta0 = S.tmem + (uint32_t)(U::RCOLS * ri - U::SHIFTC + 16 * j) + ((uint32_t)(32 * q) << 16);
    // Tiles are not separated by CTA barriers: a warp that has read its columns out of a TMEM buffer arrives on the buffer's "empty"
    // mbarrier and goes on folding; one thread of warp 0 alone waits for the arrivals, issues the MMAs of the tile after next into the buffer,
    // then folds its own columns.  Every mbarrier completes an even number of times per multiplication (four tiles per phase), so
    // all wait parities are constants.
    {
        // HIGH, issue slot s is tile t = NT_H - 1 - s: top digit of this warp's range d_top = DT (s + 1) - 1 - RD ri
        const int d_top0 = U::DT - 1 - U::RD * ri;
        const int r = (d_top0 - 2) & 3;
        unsigned char* dst0 = S.asc() + U::FRONT * U::CHB + (ge * 32 + lane) * 16 + (((d_top0 - 2) >> 2) * U::CHB);      // (d_top0 - 2) >> 2 may be -1: floor
#pragma unroll
        for (int s = 0; s < U::NT_H; s++) {
            mbar_wait(&S.bars()[s & 1], U::par_full(true, s), dead);
            tc_fence_after();
            int va[4 * (U::RD + 1)];
            umma_load<U::RD>(ta0 + (uint32_t)((s & 1) * U::TN), va);
            if (s + 2 < U::NT_H) {
                tc_fence_before();
                if (lane == 0) mbar_arrive(&S.bars()[2 + (s & 1)]);
                __syncwarp();
                if (cwarp == 0) {
                    if (EL ? elect_one() : lane == 0) {
                        mbar_wait(&S.bars()[2 + (s & 1)], U::par_empty(true, s), dead);
                        tc_fence_after();
                        umma_issue<C, LG, WIT, true>(S, s + 2);
                    }
                    __syncwarp();
                }
            }
            umma_fold<C, U::CHB, U::RD, true>(va, dst0 + s * ((U::DT / 4) * U::CHB), d_top0 + U::DT * s, r);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();                    // q-hat rows complete
        if (cwarp == 0) {
            if (EL ? elect_one() : lane == 0) { tc_fence_after(); umma_issue<C, LG, WIT, false>(S, 0); umma_issue<C, LG, WIT, false>(S, 1); }
            __syncwarp();
        }
    }
    {
        // LOW, issue slot s is tile t = s: d_top = DT (NT_L - s) - 1 - RD ri;  F of ciphertext group ge is that group's B buffer
        const int d_top0 = U::DT * U::NT_L - 1 - U::RD * ri;
        unsigned char* dst0 = S.base + (2 * ge + 1) * U::VALB + lane * 4 + d_top0 * 128;
#pragma unroll
        for (int s = 0; s < U::NT_L; s++) {
            mbar_wait(&S.bars()[s & 1], U::par_full(false, s), dead);
            tc_fence_after();
            int va[4 * (U::RD + 1)];
            umma_load<U::RD>(ta0 + (uint32_t)((s & 1) * U::TN), va);
            if (s + 2 < U::NT_L) {
                tc_fence_before();
                if (lane == 0) mbar_arrive(&S.bars()[2 + (s & 1)]);
                __syncwarp();
                if (cwarp == 0) {
                    if (EL ? elect_one() : lane == 0) {
                        mbar_wait(&S.bars()[2 + (s & 1)], U::par_empty(false, s), dead);
                        tc_fence_after();
                        umma_issue<C, LG, WIT, false>(S, s + 2);
                    }
                    __syncwarp();
                }
            }
            umma_fold<C, U::CHB, U::RD, false>(va, dst0 - s * (U::DT * 128), d_top0 - U::DT * s, 0);
        }
        tc_fence_before();
        __syncthreads();                    // F complete, TMEM free for the next phase A's stash
    }
    if (WIT) return;
    // V block = ripple(T block - F), carry into digit 0 of the next block (all MMAs are complete: the q1 rows over V are dead)
    {
        int a[C::CH * 4];
        load_block<C>(a, S.tblk(warp, lane), 32);
        const int* F = (const int*)S.B + (warp * BL) * 32 + lane;
        int carry = 0;
#pragma unroll
        for (int k = 0; k < BL; k++) {
            const int tt = (a[k] - F[k * 32] + (1 << (W - 1))) + carry;
            carry = tt >> W;
            a[k] = (tt & ((1 << W) - 1)) - (1 << (W - 1));
        }
#pragma unroll
        for (int k = BL; k < C::CH * 4; k++) a[k] = 0;
        store_block<C>(blk_ptr<C>(S.V, warp, lane), a);
        __syncthreads();
        if (warp + 1 < G) *(int*)blk_ptr<C>(S.V, warp + 1, lane) += carry;
        __syncthreads();
    }
}

template <class C, int LG, bool SQR, bool EL = true>
__device__ __forceinline__ void mulmod_u(SmemU<C, LG>& S, const int4* Y) {
    phase_product_u<C, LG>((int4*)S.base, Y, SQR ? 1 : 0, S.tmem);
    phases_bc_umma<C, LG, false, EL>((int4*)S.base, S.tmem);
}

// ---- witness step on the tcgen05 phases --------------------------------------------------------------------------------------
// w_tail of block28.cuh on this engine's buffers: the raw digits of V' are T_lo - F, q-hat is decoded from the rows of phase C,
// the carry / comparison flags live in B (F is dead once every thread holds its digits), the per-block hash sums in the gap
// behind B, the flat canonical digits over T as before.  Same arithmetic, same records.
template <class C>
__device__ __noinline__ u64w w_tail_u(int4* smem_base, int4* next_dst, WStep out, int sh, int words_out,
                                      double inv, const u64w* __restrict__ cpow) {
    typedef UL<C, 1, true> U;
    constexpr int G = C::G, BL = C::BL, L = C::L, MSK = (1 << W) - 1;
    SmemU<C, 1, true> S(smem_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* est = (int*)(S.base + U::OFF_EST);
    int* scr = (int*)S.B;
    int* cfR[2] = {scr, scr + G * 32};
    int* cfQ[2] = {scr + 2 * G * 32, scr + 3 * G * 32};
    int* cmpf = scr + 4 * G * 32;
    u64w* hsum = (u64w*)(S.base + 2 * U::VALB);
    const int* ntu = (const int*)S.two_sh + warp * C::CH * 4;      // unsigned digits of Nt_w, this block (broadcast reads)
    int rd[BL], qd[BL];
    {
        int a[C::CH * 4];
        load_block<C>(a, blk_ptr<C>(S.T, warp, lane), 32);
        const int* F = (const int*)S.B + (warp * BL) * 32 + lane;
        const unsigned char* row = S.asc() + U::FRONT * U::CHB + lane * 16;
#pragma unroll
        for (int k = 0; k < BL; k++) {
            const int j = warp * BL + k;
            rd[k] = a[k] - F[k * 32];
            const unsigned v = *(const unsigned*)(row + (j >> 2) * U::CHB + (j & 3) * 4);
            qd[k] = ((int)(v << 24) >> 24) + (((int)(v << 16) >> 24) << 7) + (((int)(v << 8) >> 24) << 14) + (((int)v >> 24) << 21);
        }
    }
    if (warp == G - 1) {       // k estimate: strict top two digits of V' (computed mod 2^(28L), |V'| < 2^beta)
        int carry = 0, d_hi = 0, d_lo = 0;
#pragma unroll
        for (int k = 0; k < BL; k++) {
            int t = rd[k] + carry + (1 << (W - 1));
            carry = t >> W;
            int d = (t & MSK) - (1 << (W - 1));
            if (k == BL - 2) d_lo = d;
            if (k == BL - 1) d_hi = d;
        }
        double vt = (double)d_hi * 268435456.0 + (double)d_lo;
        int ke = (int)floor(vt * inv);
        est[lane] = ke < -3 ? -3 : (ke > 3 ? 3 : ke);
    }
    __syncthreads();
    int adj = est[lane];
    int rnd = 0;
    bool first = true;
    for (;;) {
        int cR = 0, cQ = (warp == 0) ? adj : 0;
        if (first || adj) {
#pragma unroll
            for (int k = 0; k < BL; k++) { int t = rd[k] - adj * ntu[k] + cR; rd[k] = t & MSK; cR = t >> W; }
#pragma unroll
            for (int k = 0; k < BL; k++) { int t = qd[k] + cQ; qd[k] = t & MSK; cQ = t >> W; }
        } else cQ = 0;
        first = false;
        for (;;) {
            if (warp == G - 1) { cR = 0; cQ = 0; }
            cfR[rnd][warp * 32 + lane] = cR; cfQ[rnd][warp * 32 + lane] = cQ;
            if (!__syncthreads_or((cR | cQ) != 0)) break;
            const int iR = warp ? cfR[rnd][(warp - 1) * 32 + lane] : 0, iQ = warp ? cfQ[rnd][(warp - 1) * 32 + lane] : 0;
            cR = add_carry_block<BL>(rd, iR);
            cQ = add_carry_block<BL>(qd, iQ);
            rnd ^= 1;
        }
        int cmp = 0;
#pragma unroll
        for (int k = 0; k < BL; k++) if (rd[k] != ntu[k]) cmp = rd[k] > ntu[k] ? 1 : -1;
        if (warp == G - 1 && rd[BL - 1] >= (1 << (W - 1))) cmp = -2;
        cmpf[warp * 32 + lane] = cmp;
        __syncthreads();
        int c = 0;
#pragma unroll
        for (int b = G - 1; b >= 0; b--) { const int f = cmpf[b * 32 + lane]; if (c == 0) c = f; }
        adj = c == -2 ? -1 : (c >= 0 ? 1 : 0);
        if (!__syncthreads_or(adj != 0)) break;
        rnd ^= 1;
    }
    int* Rf = (int*)S.T;
    int* Qf = Rf + L * 32;
#pragma unroll
    for (int k = 0; k < BL; k++) { Rf[(warp * BL + k) * 32 + lane] = rd[k]; Qf[(warp * BL + k) * 32 + lane] = qd[k]; }
    __syncthreads();
    u64w h = 0;
    const int wpw = (words_out + G - 1) / G;
    for (int i = 0; i < wpw; i++) {
        const int j = warp * wpw + i;
        if (j < words_out) {
            const u64w wq = extract64<L>(Qf, 64 * j, lane), wr = extract64<L>(Rf, 64 * j + sh, lane);
            h += wq * cpow[j] + wr * cpow[words_out + j];
            if (out.rec) { out.rec[j] = wq; out.rec[words_out + j] = wr; }
            if (out.rem_out) out.rem_out[j] = wr;
            if (out.q_out) out.q_out[j] = wq;
        }
    }
    hsum[warp * 32 + lane] = h;
    int carry = 0;
    if (next_dst) {            // strict digits of R >> s (= rem 2^s), this block
        const int s = sh >> 1, pd = s / W, off = s - pd * W;
        int a[C::CH * 4];
#pragma unroll
        for (int k = 0; k < BL; k++) {
            const int p = warp * BL + k + pd;
            const unsigned lo = p < L ? (unsigned)Rf[p * 32 + lane] : 0u, hi = p + 1 < L ? (unsigned)Rf[(p + 1) * 32 + lane] : 0u;
            const int x = (int)(((lo >> off) | (off ? hi << (W - off) : 0u)) & MSK);
            const int t = x + carry + (1 << (W - 1));
            carry = t >> W;
            a[k] = (t & MSK) - (1 << (W - 1));
        }
#pragma unroll
        for (int k = BL; k < C::CH * 4; k++) a[k] = 0;
        store_block<C>(blk_ptr<C>(next_dst, warp, lane), a);
    }
    __syncthreads();
    if (next_dst && warp + 1 < G) *(int*)blk_ptr<C>(next_dst, warp + 1, lane) += carry;
    u64w H = 0;
    if (warp == 0) {
#pragma unroll
        for (int b = 0; b < G; b++) H += hsum[b * 32 + lane];
    }
    __syncthreads();
    return H;
}

// one witnessed mul_mod on the tcgen05 phases; same contract as mulmod_w (block28.cuh): V is preserved, the hash is valid in warp 0
template <class C>
__device__ __forceinline__ u64w mulmod_wu(int4* smem_base, uint32_t tmem, const int4* Y, int sqr, int4* next_dst, WStep out,
                                          int sh, int words_out, double inv, const u64w* cpow) {
    phase_product_u<C, 1, true>(smem_base, Y, sqr, tmem);
    phases_bc_umma<C, 1, true>(smem_base, tmem);
    return w_tail_u<C>(smem_base, next_dst, out, sh, words_out, inv, cpow);
}

// host: CM[u][r][b] = K7c[z0 - (8u + r + b)] (zero outside the constant)
template <class C, int LG>
inline void umma_cm_table(const signed char* k7, bool high, signed char* out /* ncm * 128 */) {
    using U = UL<C, LG>;
    const int z0 = high ? U::Z0_H : U::Z0_L, n = high ? U::NCM_H : U::NCM_L;
    for (int u = 0; u < n; u++)
        for (int r = 0; r < 8; r++)
            for (int b = 0; b < 16; b++) {
                const int d = z0 - (8 * u + r + b);
                out[(u * 8 + r) * 16 + b] = (d >= 0 && d < U::K7) ? k7[d] : (signed char)0;
            }
}

}  // namespace b28
}  // namespace pb200
