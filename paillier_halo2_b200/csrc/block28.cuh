// block28.cuh — "block28" engine: warp-role block products in radix 2^28 with lazy Barrett.
//
// Why this shape (measured on B200, profiles/imad_peak_r01.json): IMAD.WIDE without carry runs at
// 63.5 MAC/clk/SM; the carry-chained IMAD.WIDE.X form of 32-bit-limb arithmetic runs at 31.3.  So the
// hot loop uses signed 28-bit digits and plain 64-bit column accumulators (mad.wide.s32, no carry
// flags): up to 304 products of |d| <= 2^27 fit a signed 64-bit column.
//
// Layout.  One CTA = G warps x 32 lanes.  A LANE owns one ciphertext; a WARP owns a ROLE.  A number is
// L = G*BL digits = G blocks of BL digits; in shared memory a block is CH = ceil(BL/4) int4 chunks,
// stored [block][chunk][lane] so a warp's access is 512 contiguous bytes (conflict-free) and every
// lane sees only its own ciphertext.  Per-key constants (mu, Nt) are stored once per CTA and read as
// broadcasts.
//
// mulmod(V, B):  all arithmetic modulo Nt = n^2 << sh (a multiple of n^2, bit length beta - MARGIN)
//   A  T  = V * B                 (2G blocks)  role t sums the block pairs of anti-diagonals t and t+G
//   B  Q  = hi(q1 * mu)           (G blocks)   q1 = T digits [L-1, 2L-1);  anti-diagonals >= G-1
//   C  V' = lo(T) - lo(Q * Nt)    (G blocks)   anti-diagonals <= G-1;  V' == V*B (mod Nt), |V'| < 2^beta
// with 2*beta = 28(2L-1).  Values stay lazily reduced (signed, |v| < 2^beta) along the whole chain;
// one exact canonicalisation happens per output (finalize).  Each anti-diagonal sum D (2BL-1 columns)
// is rippled to 2BL strict digits + a spill digit in registers, then merged in three barrier-separated
// steps: Lo digits -> block d (store), Hi digits -> block d+1 (add + per-block ripple), carry + spill
// -> digit 0 of block d+2.  tests/model_block28.py is the limb-exact Python model of this file.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pb200 {
namespace b28 {

typedef uint64_t u64w;

constexpr int W = 28;
constexpr int MARGIN = 10;

template <int G_, int BL_>
struct Cfg {
    static constexpr int G = G_, BL = BL_, L = G_ * BL_;
    static constexpr int CH = (BL_ + 3) / 4;            // int4 chunks per block
    static constexpr int BLK4 = CH * 32;                // int4 per block per CTA (32 lanes)
    static constexpr int VAL4 = G_ * BLK4;              // int4 per L-digit value per CTA
    static constexpr int NCOL = 2 * BL_ - 1;
    static constexpr int BETA = 14 * (2 * L - 1);
    static constexpr int KN = BETA - MARGIN;            // bit length of Nt
    static constexpr int THREADS = 32 * G_;
    static constexpr int ENTRY4 = G_ * CH;              // int4 per value in global tables (one lane)
    // IMMA variant of the constant-operand phases: numbers as signed 7-bit digits (4 per 28-bit digit)
    static constexpr int K7 = 4 * L;                                  // s8 digits per number
    static constexpr int KSTEPS = (K7 + 31) / 32;                     // k-steps of mma.m16n8k32
    static constexpr bool HALF_LAST = (K7 % 32) != 0;                 // last k-step only half populated (K7 % 32 == 16)
    static constexpr int RS = K7 + ((((K7 / 4) % 8) == 4) ? 0 : ((4 - ((K7 / 4) % 8) + 8) % 8) * 4);   // row stride, (RS/4) % 8 == 4
    static constexpr int PAD7 = 64;                                   // zero padding either side of the reversed constant: a quad of
                                                                      // tiles reaches 62 bytes beyond the band at both ends
    static constexpr int XLEN = K7 + 2 * PAD7;                        // reversed constant table, zero padded
    // stride between the four byte-shifted copies: XSTR / 4 == 16 (mod 32), so that the two shift classes a warp reads in one
    // instruction (even and odd MMA columns) fall on disjoint banks
    static constexpr int XSTR = ((XLEN - 64 + 127) / 128) * 128 + 64;
    static constexpr int RTAB4 = (4 * XSTR + 15) / 16;                // int4 per constant (4 byte-shifted copies)
    static constexpr int NP_HIGH = (L + 2 + 3) / 4, NP_LOW = (L + 3) / 4;   // tile pairs (4 digits each) per phase
    // shared memory: V, B, Q (L digits each), T (2L digits), constants mu, Nt, two_sh, reversed s8 tables of mu, Nt
    static constexpr int SMEM_INT4 = 5 * VAL4 + 3 * ENTRY4 + 2 * RTAB4;
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_INT4 * 16;
    static constexpr size_t SMEM_W_BYTES = SMEM_BYTES + 128;          // witness kernel: + one int per lane (k estimate)
    // resident CTAs per SM the kernels are compiled for (shared memory and 64K registers / 128 per thread)
    static constexpr int BY_SMEM = (int)((227 * 1024) / SMEM_BYTES), BY_REGS = 512 / THREADS;
    static constexpr int CTAS_PER_SM = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
};

__device__ __forceinline__ int sgxt28(int x) {
    int r; asm("bfe.s32 %0, %1, 0, 28;" : "=r"(r) : "r"(x)); return r;
}
// acc += a * b (signed 32 x 32 -> 64).  Written as a carry pair on the two halves of the accumulator, which ptxas turns into ONE
// IMAD.WIDE Rd, Ra, Rb, Rd with a 64-bit register addend; the plain `mad.wide.s32 d, a, b, d` is split by ptxas for sm_100a
// into IMAD.WIDE .., RZ plus IADD3 / IADD3.X (two instructions per product instead of one).
__device__ __forceinline__ void madw(long long& acc, int a, int b) {
#ifndef PB200_SPLIT_MAC
    asm("{ .reg .b32 l, h; mov.b64 {l, h}, %0; mad.lo.cc.s32 l, %1, %2, l; madc.hi.s32 h, %1, %2, h; mov.b64 %0, {l, h}; }"
        : "+l"(acc) : "r"(a), "r"(b));
#else
    asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#endif
}

enum Mode { M_FULL = 0, M_LT = 1, M_UTG = 2, M_SQ = 3 };

template <int BL, int MODE>
__device__ __forceinline__ constexpr bool include_xy(int x, int y) {
    return MODE == M_FULL ? true : MODE == M_LT ? (x + y <= BL - 1) : MODE == M_UTG ? (x + y >= BL - 2) : (x <= y);
}

// load one block (BL digits) of a per-lane buffer into registers.  p points at chunk 0 of the block for this lane.
template <class C>
__device__ __forceinline__ void load_block(int (&a)[C::CH * 4], const int4* p, int stride) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) {
        int4 v = p[c * stride];
        a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w;
    }
}
template <class C>
__device__ __forceinline__ void store_block(int4* p, const int (&a)[C::CH * 4]) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) p[c * 32] = make_int4(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
}

// acc[x+y] += a[x] * b[y] over the (x, y) of MODE; b streamed from shared memory chunk by chunk.
// DBL doubles every product (off-diagonal block pairs of a squaring); M_SQ doubles x<y and keeps x==y.
template <class C, int MODE, bool DBL>
__device__ __forceinline__ void mac_block(long long (&acc)[C::NCOL], const int (&a)[C::CH * 4], const int4* bp, int bstride) {
    constexpr int BL = C::BL;
#pragma unroll
    for (int c = 0; c < C::CH; c++) {
        int4 bv = bp[c * bstride];
        int b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int y = 4 * c + e;
            if (y < BL) {
                int by = b4[e];
                int by2 = by + by;
#pragma unroll
                for (int x = 0; x < BL; x++) {
                    if (include_xy<BL, MODE>(x, y)) {
                        if (MODE == M_SQ) madw(acc[x + y], a[x], x == y ? by : by2);
                        else madw(acc[x + y], a[x], DBL ? by2 : by);
                    }
                }
            }
        }
    }
}

// ripple 2BL-1 column sums into 2BL strict digits + spill.  lo[] <- digits [0,BL), hi[] <- [BL,2BL), returns spill
template <class C>
__device__ __forceinline__ int normalize(const long long (&acc)[C::NCOL], int (&lo)[C::CH * 4], int (&hi)[C::CH * 4]) {
    constexpr int BL = C::BL;
    long long carry = 0;
#pragma unroll
    for (int k = 0; k < C::NCOL; k++) {
        long long t = acc[k] + carry;
        int d = sgxt28((int)t);
        carry = (t - d) >> W;
        if (k < BL) lo[k] = d; else hi[k - BL] = d;
    }
    int c32 = (int)carry;            // |carry| < 2^35 before the last digit; after sgxt the rest is tiny
    int d = sgxt28(c32);
    hi[BL - 1] = d;
    int spill = (int)((carry - d) >> W);
#pragma unroll
    for (int k = BL; k < C::CH * 4; k++) { lo[k] = 0; hi[k] = 0; }
    return spill;
}

template <class C>
__device__ __forceinline__ void zero_acc(long long (&acc)[C::NCOL]) {
#pragma unroll
    for (int k = 0; k < C::NCOL; k++) acc[k] = 0;
}

// dst block <- ripple(src block + sign*h), returns the carry out of the block.  p_src/p_dst: chunk 0 for this lane.
template <class C>
__device__ __forceinline__ int add_ripple_block(int4* p_dst, const int4* p_src, const int (&h)[C::CH * 4], int sign) {
    constexpr int BL = C::BL;
    int a[C::CH * 4];
    load_block<C>(a, p_src, 32);
    // biased form: t' = value + 2^27 + carry; carry' = t' >> 28 (floor), digit = (t' & M) - 2^27: the serial chain
    // per digit is one add and one shift, the digit itself is off the critical path
    int carry = 0;
#pragma unroll
    for (int k = 0; k < BL; k++) {
        int s0 = a[k] + sign * h[k] + (1 << (W - 1));
        int t = s0 + carry;
        carry = t >> W;
        a[k] = (t & ((1 << W) - 1)) - (1 << (W - 1));
    }
#pragma unroll
    for (int k = BL; k < C::CH * 4; k++) a[k] = 0;
    store_block<C>(p_dst, a);
    return carry;
}

// One anti-diagonal result waiting for its barrier-separated merge steps.
template <class C>
struct Pending {
    int hi[C::CH * 4];
    int spill;
    int carry;
    int blk;      // output block index of the Lo part; -1 = nothing pending
};

// ---------------------------------------------------------------------------------------------
// Shared-memory view of one CTA
template <class C>
struct Smem {
    int4* V; int4* B; int4* Q; int4* T; const int4* mu; const int4* Nt; const int4* two_sh;   // V must stay first
    const char* rmu; const char* rnt;      // reversed, byte-shifted s8 tables of mu and Nt (IMMA phases)
    __device__ __forceinline__ Smem(int4* base) {
        V = base; B = V + C::VAL4; Q = B + C::VAL4; T = Q + C::VAL4;
        int4* k = T + 2 * C::VAL4;
        mu = k; Nt = k + C::ENTRY4; two_sh = k + 2 * C::ENTRY4;
        rmu = (const char*)(k + 3 * C::ENTRY4); rnt = (const char*)(k + 3 * C::ENTRY4 + C::RTAB4);
    }
};

template <class C>
__device__ __forceinline__ int4* blk_ptr(int4* buf, int blk, int lane) { return buf + blk * C::BLK4 + lane; }
template <class C>
__device__ __forceinline__ const int4* blk_ptr(const int4* buf, int blk, int lane) { return buf + blk * C::BLK4 + lane; }

// ---- generic phase executor ---------------------------------------------------------------------
// One non-inlined function runs all three phases of a mulmod; everything that differs between phases and
// roles is a warp-uniform runtime value, so the unrolled block-product bodies exist once per job slot.
enum Phase { PH_MUL = 0, PH_SQR = 1, PH_HIGH = 2, PH_LOW = 3 };

// load block i of q1 = T digits [L-1, 2L-1): digit 0 from the top of T block G+i-1, the rest from block G+i
template <class C>
__device__ __forceinline__ void load_q1_block(int (&a)[C::CH * 4], const int4* T, int i, int lane) {
    constexpr int G = C::G, BL = C::BL;
    int t[C::CH * 4];
    load_block<C>(t, blk_ptr<C>(T, G + i, lane), 32);
    const int* below = (const int*)blk_ptr<C>(T, G + i - 1, lane);
    a[0] = below[((BL - 1) / 4) * 32 * 4 + ((BL - 1) % 4)];
#pragma unroll
    for (int k = 1; k < BL; k++) a[k] = t[k - 1];
    if (i == G - 1) a[BL - 1] += t[BL - 1] << W;       // fold digit 2L-1 (|.| <= 1) into digit 2L-2
#pragma unroll
    for (int k = BL; k < C::CH * 4; k++) a[k] = 0;
}

// One anti-diagonal job: accumulate its block pairs, ripple, write the Lo digits; Hi digits + spill stay in registers.
template <class C>
__device__ __forceinline__ void run_job(Pending<C>& pend, int ph, int d, int i_lo, int i_hi,
                                        const int4* abuf, const int4* bbase, int bblk, int bchunk,
                                        int4* lo_dst, int lo_blk, int lo_op, int lane) {
    long long acc[C::NCOL];
    zero_acc<C>(acc);
#pragma unroll 1
    for (int i = i_lo; i <= i_hi; i++) {
        int a[C::CH * 4];
        if (ph == PH_HIGH) load_q1_block<C>(a, abuf, i, lane);
        else load_block<C>(a, blk_ptr<C>(abuf, i, lane), 32);
        const int4* bp = bbase + (d - i) * bblk;
        if (ph == PH_SQR && 2 * i == d) {
            mac_block<C, M_SQ, false>(acc, a, bp, bchunk);
        } else {
            if (ph == PH_SQR) {
#pragma unroll
                for (int k = 0; k < C::BL; k++) a[k] += a[k];      // off-diagonal pair counted twice
            }
            mac_block<C, M_FULL, false>(acc, a, bp, bchunk);
        }
    }
    int lo[C::CH * 4];
    pend.spill = normalize<C>(acc, lo, pend.hi);
    pend.carry = 0;
    if (lo_op == 1) {                    // plain store of the Lo digits
        store_block<C>(blk_ptr<C>(lo_dst, lo_blk, lane), lo);
    } else if (lo_op == 2) {             // subtract from the block in place (phase C)
        int4* p = blk_ptr<C>(lo_dst, lo_blk, lane);
        int t[C::CH * 4];
        load_block<C>(t, p, 32);
#pragma unroll
        for (int k = 0; k < C::BL; k++) t[k] -= lo[k];
        store_block<C>(p, t);
    }
}

template <class C>
__device__ __forceinline__ void store_zero_block(int4* buf, int blk, int lane) {
    int z[C::CH * 4];
#pragma unroll
    for (int k = 0; k < C::CH * 4; k++) z[k] = 0;
    store_block<C>(blk_ptr<C>(buf, blk, lane), z);
}

// Which anti-diagonal a warp works on in the truncated phases.  Jobs sorted by size (G, G-1, .., 1 block
// pairs); warps w and w + G/2 .. share a scheduler (warp id mod 4), so big jobs are paired with small ones
// per scheduler in each phase, and each warp's sizes over phases B and C add up to G + 1.
template <class C>
__device__ __forceinline__ int job_rank(int warp) {     // 0 = largest job
    return warp < C::G / 2 ? warp : (C::G - 1) - (warp - C::G / 2);
}

// ph = PH_MUL / PH_SQR : T  = V * Y          (Y = V for PH_SQR)
// ph = PH_HIGH         : Q  = digits [L,2L) of q1(T) * mu
// ph = PH_LOW          : V  = lo_L(T) - lo_L(Q * Nt), rippled
template <class C>
__device__ __noinline__ void run_phase(int4* smem_base, const int4* Y, int ph) {
    constexpr int G = C::G;
    Smem<C> S(smem_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Pending<C> pend;
    pend.blk = -1;
    int blk0 = -1, spill0 = 0;                      // first job of a two-job phase (its Hi digits are stashed in Q)
    int4* stash = S.Q + threadIdx.x;                // [chunk][thread], free during phase A
    // per-phase parameters (all warp-uniform); run_job has ONE call site so its unrolled bodies exist once
    int njobs = 1, nblk = G, sign = 1, bblk = C::CH, bchunk = 1, lo_op = 1;
    const int4* abuf = S.V; const int4* bbase = S.mu; int4* lo_dst = S.Q;
    int4* hi_dst = S.Q; const int4* hi_src = S.Q;
    int d_first = 0;
    if (ph <= PH_SQR) {
        njobs = warp < G - 1 ? 2 : 1;
        abuf = S.V; bbase = (ph == PH_SQR ? S.V : Y) + lane; bblk = C::BLK4; bchunk = 32;
        lo_dst = S.T; hi_dst = S.T; hi_src = S.T; nblk = 2 * G; d_first = warp;
    } else if (ph == PH_HIGH) {
        abuf = S.T; bbase = S.mu; d_first = G - 1 + job_rank<C>(warp);      // anti-diagonals G-1 .. 2G-2
    } else {
        abuf = S.Q; bbase = S.Nt; lo_dst = S.T; lo_op = 2; d_first = G - 1 - job_rank<C>(warp);   // G-1 .. 0, reversed pairing
        hi_dst = S.V; hi_src = S.T; sign = -1;
    }
#pragma unroll 1
    for (int half = 0; half < njobs; half++) {
        const int d = d_first + half * G;
        int i_lo, i_hi, lo_blk, op = lo_op;
        if (ph <= PH_SQR) {
            i_lo = half ? d - G + 1 : 0;
            i_hi = ph == PH_SQR ? d / 2 : (half ? G - 1 : d);
            lo_blk = d;
        } else if (ph == PH_HIGH) {
            i_lo = d >= G ? d - G + 1 : 0; i_hi = G - 1;
            lo_blk = d - G;
            if (d < G) { op = 0; lo_blk = 0; }      // anti-diagonal G-1: its Lo part lies below digit L
        } else {
            i_lo = 0; i_hi = d; lo_blk = d;
        }
        run_job<C>(pend, ph, d, i_lo, i_hi, abuf, bbase, bblk, bchunk, lo_dst, lo_blk, op, lane);
        pend.blk = (ph == PH_HIGH && d < G) ? -1 : lo_blk;
        if (half == 0 && njobs == 2) {
#pragma unroll
            for (int c = 0; c < C::CH; c++)
                stash[c * C::THREADS] = make_int4(pend.hi[4 * c], pend.hi[4 * c + 1], pend.hi[4 * c + 2], pend.hi[4 * c + 3]);
            blk0 = d; spill0 = pend.spill;
        }
    }
    if (ph <= PH_SQR) {
        if (warp == G - 1) store_zero_block<C>(S.T, 2 * G - 1, lane);       // block 2G-1 has no Lo contribution
    } else if (ph == PH_HIGH) {
        if (pend.blk < 0) store_zero_block<C>(S.Q, G - 1, lane);            // Q block G-1 has no Lo contribution
    } else if (pend.blk == G - 1) {    // no Hi below digit L: this warp ripples block 0 instead (adds zero)
#pragma unroll
        for (int k = 0; k < C::CH * 4; k++) pend.hi[k] = 0;
        pend.spill = 0; pend.blk = -1;
    }
    __syncthreads();
    // step 2: Hi digits into block blk+1 with a per-block ripple (phase C also moves the block from T to V)
    int carry0 = 0;
    if (blk0 >= 0) {
        int h0[C::CH * 4];
        load_block<C>(h0, stash, C::THREADS);
        carry0 = add_ripple_block<C>(blk_ptr<C>(hi_dst, blk0 + 1, lane), blk_ptr<C>(hi_src, blk0 + 1, lane), h0, sign);
    }
    if (pend.blk + 1 <= nblk - 1)
        pend.carry = add_ripple_block<C>(blk_ptr<C>(hi_dst, pend.blk + 1, lane), blk_ptr<C>(hi_src, pend.blk + 1, lane), pend.hi, sign);
    __syncthreads();
    // step 3: carry of that ripple + the spill digit into digit 0 of block blk+2
    if (blk0 >= 0 && blk0 + 2 <= nblk - 1) *(int*)blk_ptr<C>(hi_dst, blk0 + 2, lane) += carry0 + sign * spill0;
    if (pend.blk + 2 <= nblk - 1) *(int*)blk_ptr<C>(hi_dst, pend.blk + 2, lane) += pend.carry + sign * pend.spill;
    __syncthreads();
}

// ---- phase A specialised for the tensor engine ---------------------------------------------------------
// Same arithmetic as run_phase<PG=0>, but with nothing else alive during the product loop (32-bit shared
// addresses, compile-time strides, merge parameters recomputed afterwards): the register allocator then has
// room to keep several IMAD.WIDE products in flight instead of stalling each IADD3 on the multiply before it.
__device__ __forceinline__ int4 lds128(unsigned addr) {
    int4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
// one triangular half of a block product: HI = false: columns x+y <= BL-1 into acc[x+y];
// HI = true: columns x+y >= BL into acc[x+y-BL].  SQ: only x <= y, off-diagonal products doubled.
template <class C, bool HI, bool SQ>
__device__ __forceinline__ void mac_half(long long (&acc)[C::BL], const int (&a)[C::CH * 4], unsigned baddr) {
    constexpr int BL = C::BL;
#pragma unroll
    for (int c = 0; c < C::CH; c++) {
        int4 bv = lds128(baddr + c * 512);
        int b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int y = 4 * c + e;
            if (y < BL) {
                int by = b4[e];
                int by2 = by + by;
#pragma unroll
                for (int x = 0; x < BL; x++) {
                    const bool in_half = HI ? (x + y >= BL) : (x + y <= BL - 1);
                    if (in_half && (!SQ || x <= y)) madw(acc[HI ? x + y - BL : x + y], a[x], (SQ && x != y) ? by2 : by);
                }
            }
        }
    }
}

// all block pairs of anti-diagonal d, one triangular half
template <class C, bool HI>
__device__ __forceinline__ void job_half(long long (&acc)[C::BL], int d, int i_lo, int i_hi, int sqr, unsigned v_addr, unsigned y_addr) {
    unsigned aa = v_addr + i_lo * (C::BLK4 * 16), ba = y_addr + (d - i_lo) * (C::BLK4 * 16);
#pragma unroll 1
    for (int i = i_lo; i <= i_hi; i++, aa += C::BLK4 * 16, ba -= C::BLK4 * 16) {
        int a[C::CH * 4];
#pragma unroll
        for (int c = 0; c < C::CH; c++) { int4 v = lds128(aa + c * 512); a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w; }
        if (sqr && 2 * i == d) {
            mac_half<C, HI, true>(acc, a, ba);
        } else {
            if (sqr) {
#pragma unroll
                for (int k = 0; k < C::BL; k++) a[k] += a[k];
            }
            mac_half<C, HI, false>(acc, a, ba);
        }
    }
}

// ripple BL column sums (+ carry in) into BL strict digits; returns the 64-bit carry out
template <class C>
__device__ __forceinline__ long long ripple_cols(const long long (&acc)[C::BL], int ncols, long long carry, int (&dig)[C::CH * 4]) {
#pragma unroll
    for (int k = 0; k < C::BL; k++) {
        if (k < ncols) {
            long long t = (acc[k] + (1ll << (W - 1))) + carry;     // biased: chain = 64-bit add -> 64-bit shift
            carry = t >> W;
            dig[k] = ((int)t & ((1 << W) - 1)) - (1 << (W - 1));
        }
    }
    return carry;
}

template <class C>
__device__ __noinline__ void phase_product(int4* smem_base, const int4* Y, int sqr) {
    constexpr int G = C::G, BL = C::BL;
    Smem<C> S(smem_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v_addr = (unsigned)__cvta_generic_to_shared(S.V) + lane * 16;
    const unsigned y_addr = sqr ? v_addr : (unsigned)__cvta_generic_to_shared(Y) + lane * 16;
    Pending<C> pend;
    pend.blk = -1;
    int blk0 = -1, spill0 = 0;
    int4* stash = S.Q + threadIdx.x;
    const int njobs = warp < G - 1 ? 2 : 1;
#pragma unroll 1
    for (int half = 0; half < njobs; half++) {
        const int d = warp + half * G;
        const int i_lo = half ? d - G + 1 : 0;
        const int i_hi = sqr ? d / 2 : (half ? G - 1 : d);
#ifdef PB200_PHASEA_SINGLE
        {   // one pass over the block pairs, all 2BL-1 columns live (A/B variant)
            long long acc[C::NCOL];
            zero_acc<C>(acc);
            unsigned aa = v_addr + i_lo * (C::BLK4 * 16), ba = y_addr + (d - i_lo) * (C::BLK4 * 16);
#pragma unroll 1
            for (int i = i_lo; i <= i_hi; i++, aa += C::BLK4 * 16, ba -= C::BLK4 * 16) {
                int a[C::CH * 4];
#pragma unroll
                for (int c = 0; c < C::CH; c++) { int4 v = lds128(aa + c * 512); a[4 * c] = v.x; a[4 * c + 1] = v.y; a[4 * c + 2] = v.z; a[4 * c + 3] = v.w; }
                const bool diag = sqr && 2 * i == d;
                if (sqr && !diag) {
#pragma unroll
                    for (int k = 0; k < C::BL; k++) a[k] += a[k];
                }
#pragma unroll
                for (int c = 0; c < C::CH; c++) {
                    int4 bv = lds128(ba + c * 512);
                    int b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int y = 4 * c + e;
                        if (y < BL) {
                            if (diag) {
                                const int by = b4[e], by2 = by + by;
#pragma unroll
                                for (int x = 0; x <= y; x++) madw(acc[x + y], a[x], x == y ? by : by2);
                            } else {
#pragma unroll
                                for (int x = 0; x < BL; x++) madw(acc[x + y], a[x], b4[e]);
                            }
                        }
                    }
                }
            }
            int lo[C::CH * 4];
            pend.spill = normalize<C>(acc, lo, pend.hi);
            store_block<C>(blk_ptr<C>(S.T, d, lane), lo);
        }
#else
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
            store_block<C>(blk_ptr<C>(S.T, d, lane), lo);
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
#endif
        pend.carry = 0;
        pend.blk = d;
        if (half == 0 && njobs == 2) {
#pragma unroll
            for (int c = 0; c < C::CH; c++)
                stash[c * C::THREADS] = make_int4(pend.hi[4 * c], pend.hi[4 * c + 1], pend.hi[4 * c + 2], pend.hi[4 * c + 3]);
            blk0 = d; spill0 = pend.spill;
        }
    }
    if (warp == G - 1) store_zero_block<C>(S.T, 2 * G - 1, lane);
    __syncthreads();
    int carry0 = 0;
    if (blk0 >= 0) {
        int h0[C::CH * 4];
        load_block<C>(h0, stash, C::THREADS);
        int4* p = blk_ptr<C>(S.T, blk0 + 1, lane);
        carry0 = add_ripple_block<C>(p, p, h0, 1);
    }
    {
        int4* p = blk_ptr<C>(S.T, pend.blk + 1, lane);
        pend.carry = add_ripple_block<C>(p, p, pend.hi, 1);
    }
    __syncthreads();
    if (blk0 >= 0) *(int*)blk_ptr<C>(S.T, blk0 + 2, lane) += carry0 + spill0;
    if (pend.blk + 2 <= 2 * G - 1) *(int*)blk_ptr<C>(S.T, pend.blk + 2, lane) += pend.carry + pend.spill;
    __syncthreads();
}

// ---- IMMA variant of phases B and C ------------------------------------------------------------------
// Both multiply a per-lane number by a per-key constant: over the 32 lanes of the CTA that is a GEMM
//   C[lane][p] = sum_k A7[lane][k] * K7[p - k]          (A7, K7: signed 7-bit digits, 4 per 28-bit digit)
// run on the tensor pipe with mma.sync.m16n8k32.s8 (measured 1880 MAC/clk/SM, profiles/imma_peak_r01.json,
// against 26 MAC/clk/SM for the IMAD inner loop).  The Toeplitz operand is never materialised: a B fragment is
// four consecutive bytes of the REVERSED constant, read from one of four byte-shifted copies so every read is
// an aligned 32-bit word.  The s32 columns are folded back into 28-bit digits (lo + carry into the next digit).
__device__ __forceinline__ void mma_s8(int (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int sgxt7(int x) { int r; asm("bfe.s32 %0, %1, 0, 7;" : "=r"(r) : "r"(x)); return r; }
// one 28-bit digit -> four signed 7-bit digits packed in a word (the top one absorbs the remainder, |.| <= 65)
__device__ __forceinline__ unsigned split7_pack(int d) {
    // bias every 7-bit field by 64 so that the three low fields can be cut out as unsigned bit fields, then remove the bias
    // per byte: with f in [0, 127], (f + 0x40) ^ 0x80 == f - 64 (mod 256) and f + 0x40 never carries into the next byte
    const unsigned dp = (unsigned)d + (64u | (64u << 7) | (64u << 14));
    const unsigned x = (dp & 0x7Fu) | ((dp << 1) & 0x7F00u) | ((dp << 2) & 0x7F0000u);
    return ((x + 0x404040u) ^ 0x808080u) | ((unsigned)((int)dp >> 21) << 24);
}

// A-operand bytes live in the Q buffer: As[lane][RS]; output arrays LO (in T blocks G..2G-1) and CA (in B):
// [digit][lane rotated by 8*digit] so that both the fold (lanes vary in (g, t)) and the ripple (lanes vary in
// the ciphertext index) hit 32 distinct banks.
template <class C> __device__ __forceinline__ char* as_ptr(Smem<C>& S) { return (char*)S.Q; }
template <class C> __device__ __forceinline__ int* lo_ptr(Smem<C>& S) { return (int*)(S.T + C::VAL4); }
template <class C> __device__ __forceinline__ int* ca_ptr(Smem<C>& S) { return (int*)S.B; }
__device__ __forceinline__ int dl_index(int jj, int m) { return jj * 32 + ((m + 8 * jj) & 31); }

__device__ __forceinline__ unsigned lds32(unsigned addr) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

// C[lane][p] for the tile pairs of this warp, p = p_base + 16*U + ...; folded: LO[jj] / CA[jj+1], jj = 4U + t.
// A warp works on a QUAD = two adjacent tile pairs (32 columns, 8 digits) per pass: the A fragments of a k-step serve both, and by
// the Toeplitz structure the B fragments of the second pair are those of the first shifted by half a k-step — b0(U+1, ks) is
// b1(U, ks-1) and b1(U+1, ks) is b0(U, ks) — so a k-step costs two ldmatrix.x4 and four 32-bit loads for EIGHT mma.sync
// (the one-pair-per-pass version: the same loads for four; its IMMA phases ran at 69 % shared-memory pipe and 8.5 instructions
// per MMA, profiles/ncu_k_encrypt_r01_fused_summary.txt).
template <class C, bool HIGH>
__device__ __noinline__ void phase_mma(int4* smem_base) {
    constexpr int G = C::G, L = C::L, K7 = C::K7;
    Smem<C> S(smem_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int P_BASE = HIGH ? 4 * (L - 2) : 0;
    constexpr int NP = HIGH ? C::NP_HIGH : C::NP_LOW;
    // dealing: RQ full rounds in which every warp takes a quad, then one tail round for the R < 2G pairs left: R - G warps take a
    // quad and the others a single pair (R > G), or R warps take a single pair — no warp ever holds more than ceil(NP / G) pairs
    constexpr int RQ = NP / (2 * G), R = NP - 2 * G * RQ, NROUNDS = RQ + (R > 0 ? 1 : 0), TQ = R > G ? R - G : 0;
    constexpr int NOUT = HIGH ? L + 2 : L;
    // ldmatrix.x4 row addresses: lanes 0-7 rows 0-7 (k bytes 0-15), 8-15 rows 8-15, 16-23 rows 0-7 (+16), 24-31 rows 8-15 (+16)
    const unsigned as_base = (unsigned)__cvta_generic_to_shared(as_ptr<C>(S)) + ((lane & 7) + ((lane >> 3) & 1) * 8) * C::RS + (lane >> 4) * 16;
    const unsigned rt_base = (unsigned)__cvta_generic_to_shared(HIGH ? S.rmu : S.rnt);
    int* LO = lo_ptr<C>(S);
    int* CA = ca_ptr<C>(S);
    // fold of one tile pair: thread (g, t) owns the 4 radix-2^7 columns of digit jj = 4U + t for rows g, g+8 of each m-tile:
    // v = c0 + c1 2^7 + c2 2^14 + c3 2^21 (|c| < 2^23) in 32-bit pieces, lo digit + carry into the next digit
    auto fold = [&](const int (&acc)[2][2][4], int U) {
        const int jj = 4 * U + t;
        if (jj < NOUT) {
            const int x0 = (g + 8 * jj) & 31, x1 = (g + 8 * (jj + 1)) & 31;
            int* lo_row = LO + jj * 32;
            int* ca_row = CA + (jj + 1) * 32;
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int lowp = acc[0][mt][2 * r] + (acc[0][mt][2 * r + 1] << 7);
                    const int highp = acc[1][mt][2 * r] + (acc[1][mt][2 * r + 1] << 7);
                    const int tb = lowp + ((highp & 0x3FFF) << 14) + (1 << (W - 1));       // biased low part
                    const int lo = (tb & ((1 << W) - 1)) - (1 << (W - 1));
                    const int ca = (tb >> W) + (highp >> 14);
                    const int off = 16 * mt + 8 * r;
                    lo_row[(x0 + off) & 31] = lo;
                    ca_row[(x1 + off) & 31] = ca;
                }
        }
    };
#pragma unroll 1
    for (int q = 0; q < NROUNDS; q++) {
        // serpentine order over the rounds (the k-steps of a tile pair grow (phase C) or shrink (phase B) linearly with its index and
        // the phase ends at a CTA barrier)
        const int ws = (q & 1) ? (G - 1 - warp) : warp;
        int U; bool two;
        if (q < RQ) { U = 2 * (q * G + ws); two = true; }
        else if (ws < TQ) { U = 2 * G * RQ + 2 * ws; two = true; }
        else { U = 2 * G * RQ + 2 * TQ + (ws - TQ); two = false; if (U >= NP) continue; }
        const int P0 = P_BASE + 16 * U;
        // k-steps with some (k, p): 0 <= p - k <= K7-1, k in [32ks, 32ks+32), p in [P0, P0+32)
        int ks_lo = (P0 - (K7 - 1)) / 32; if (P0 - (K7 - 1) <= 0) ks_lo = 0;
        int ks_hi = (P0 + (two ? 31 : 15)) / 32; if (ks_hi > C::KSTEPS - 1) ks_hi = C::KSTEPS - 1;
        // B-fragment addresses of the first pair: column n = g of tile h is p = P0 + 4(g>>1) + (g&1) + 2h; 4 ascending bytes of the
        // reversed table, taken from the copy whose shift makes the read an aligned word
        unsigned baddr[2], prev[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = P0 + 4 * (g >> 1) + (g & 1) + 2 * h;
            const int idx0 = C::PAD7 + K7 - 1 - p + 4 * t;
            const int sft = idx0 & 3;
            baddr[h] = rt_base + sft * C::XSTR + (idx0 - sft) + 32u * ks_lo;
            prev[h] = lds32(baddr[h] - 16);
        }
        int acc0[2][2][4], acc1[2][2][4];           // [tile h][m-tile][c0..c3] of the two pairs
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int i = 0; i < 4; i++) { acc0[h][mt][i] = 0; acc1[h][mt][i] = 0; }
        unsigned ap = as_base + 32u * ks_lo;
        const int nks = ks_hi - ks_lo + 1;
        const int half_at = (C::HALF_LAST && ks_hi == C::KSTEPS - 1) ? nks - 1 : -1;      // the half populated last k-step
        if (two) {
#pragma unroll 2
            for (int i = 0; i < nks; i++, ap += 32, baddr[0] += 32, baddr[1] += 32) {
                unsigned a[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3]) : "r"(ap + mt * 16 * C::RS));
                    if (i == half_at) { a[mt][2] = 0; a[mt][3] = 0; }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const unsigned b0 = lds32(baddr[h]), b1 = lds32(baddr[h] + 16);
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        mma_s8(acc0[h][mt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
                        mma_s8(acc1[h][mt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], prev[h], b0);
                    }
                    prev[h] = b1;
                }
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < nks; i++, ap += 32, baddr[0] += 32, baddr[1] += 32) {
                unsigned a[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3]) : "r"(ap + mt * 16 * C::RS));
                    if (i == half_at) { a[mt][2] = 0; a[mt][3] = 0; }
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const unsigned b0 = lds32(baddr[h]), b1 = lds32(baddr[h] + 16);
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) mma_s8(acc0[h][mt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
                }
            }
        }
        fold(acc0, U);
        if (two) fold(acc1, U + 1);
    }
    __syncthreads();
}

// One block of packed s8 digits (BL words) into a lane's row of As.  Rows are RS bytes apart: a 32-bit store per lane touches only
// 8 banks (4-way conflict, 76 wavefronts per block), a 128-bit store per lane is conflict-free (a quarter warp covers all 32 banks).
// The block starts at word warp*BL of the 16-byte aligned row, i.e. at one of four alignments: LEAD single words reach the next
// 16-byte boundary, the rest goes out as 128-bit stores plus a short tail.
template <int BL, int LEAD>
__device__ __forceinline__ void store_row_aligned(unsigned* row, const unsigned (&w)[BL]) {
#pragma unroll
    for (int k = 0; k < LEAD; k++) row[k] = w[k];
    constexpr int NV = (BL - LEAD) / 4;
#pragma unroll
    for (int c = 0; c < NV; c++)
        *reinterpret_cast<uint4*>(row + LEAD + 4 * c) = make_uint4(w[LEAD + 4 * c], w[LEAD + 4 * c + 1], w[LEAD + 4 * c + 2], w[LEAD + 4 * c + 3]);
#pragma unroll
    for (int k = LEAD + 4 * NV; k < BL; k++) row[k] = w[k];
}
template <int BL>
__device__ __forceinline__ void store_row(unsigned* row, const unsigned (&w)[BL], int warp) {
    switch ((4 - ((warp * BL) & 3)) & 3) {          // warp-uniform
        case 0: store_row_aligned<BL, 0>(row, w); break;
        case 1: store_row_aligned<BL, 1>(row, w); break;
        case 2: store_row_aligned<BL, 2>(row, w); break;
        default: store_row_aligned<BL, 3>(row, w); break;
    }
}

// q1 (T digits [L-1, 2L-1)) -> s8 rows in As.  One (block = warp, lane) per thread.
template <class C>
__device__ __forceinline__ void q1_to_bytes(Smem<C>& S, int warp, int lane) {
    int a[C::CH * 4];
    load_q1_block<C>(a, S.T, warp, lane);
    unsigned* row = (unsigned*)(as_ptr<C>(S) + lane * C::RS) + warp * C::BL;
    unsigned w[C::BL];
#pragma unroll
    for (int k = 0; k < C::BL; k++) w[k] = split7_pack(a[k]);
    store_row<C::BL>(row, w, warp);
    if (C::K7 < C::KSTEPS * 32 && warp == C::G - 1) {           // zero the tail of a half populated last k-step
        unsigned* tail = (unsigned*)(as_ptr<C>(S) + lane * C::RS) + C::L;
#pragma unroll
        for (int k = 0; k < (C::KSTEPS * 32 - C::K7) / 4; k++) if (4 * (C::L + k) < C::RS) tail[k] = 0;
    }
    __syncthreads();
}

// phase B tail: q-hat digits (jj = j + 2, two guard digits below) as s8 rows in As.  The digits are NOT rippled: LO + CA is within
// 2^27 + 2^17 in magnitude, which the s8 split absorbs in its top piece ([-65, 64]); q-hat is the same integer, phase C is exact on
// any digit representation and the s32 columns stay below 608 * 65 * 64.  Only the two guard digits feed a carry into digit 0.
// (tests/model_block28.py models exactly this; the rippled version cost a 19-step serial chain and two more CTA barriers.)
template <class C>
__device__ __forceinline__ void qhat_to_bytes(Smem<C>& S, int warp, int lane) {
    const int* LO = lo_ptr<C>(S);
    const int* CA = ca_ptr<C>(S);
    int carry = 0;
    if (warp == 0) {
        int t0 = LO[dl_index(0, lane)];
        carry = (t0 - sgxt28(t0)) >> W;
        int t1 = LO[dl_index(1, lane)] + CA[dl_index(1, lane)] + carry;
        carry = (t1 - sgxt28(t1)) >> W;
    }
    unsigned* row = (unsigned*)(as_ptr<C>(S) + lane * C::RS) + warp * C::BL;      // As (Q buffer) is dead since phase B's MMAs ended
    unsigned w[C::BL];
#pragma unroll
    for (int k = 0; k < C::BL; k++) {
        const int jj = warp * C::BL + k + 2;
        w[k] = split7_pack(LO[dl_index(jj, lane)] + CA[dl_index(jj, lane)] + (k == 0 ? carry : 0));
    }
    store_row<C::BL>(row, w, warp);
    __syncthreads();
}

// phase C tail: V block = ripple(T_lo block - LO - CA), carry into digit 0 of the next block
template <class C>
__device__ __forceinline__ void low_to_value(Smem<C>& S, int warp, int lane) {
    const int* LO = lo_ptr<C>(S);
    const int* CA = ca_ptr<C>(S);
    int a[C::CH * 4];
    load_block<C>(a, blk_ptr<C>(S.T, warp, lane), 32);
    int carry = 0;
#pragma unroll
    for (int k = 0; k < C::BL; k++) {
        const int j = warp * C::BL + k;
        int tt = (a[k] - LO[dl_index(j, lane)] - (j > 0 ? CA[dl_index(j, lane)] : 0) + (1 << (W - 1))) + carry;
        carry = tt >> W;
        a[k] = (tt & ((1 << W) - 1)) - (1 << (W - 1));
    }
#pragma unroll
    for (int k = C::BL; k < C::CH * 4; k++) a[k] = 0;
    store_block<C>(blk_ptr<C>(S.V, warp, lane), a);
    __syncthreads();
    if (warp + 1 < C::G) *(int*)blk_ptr<C>(S.V, warp + 1, lane) += carry;
    __syncthreads();
}

// V <- V * Y mod Nt (lazy).  SQR: Y ignored, V <- V^2.
template <class C, bool SQR, bool MMA = false>
__device__ __forceinline__ void mulmod(Smem<C>& S, const int4* Y, int role, int lane) {
    int4* base = S.V;
    if (MMA) phase_product<C>(base, Y, SQR ? 1 : 0);
    else run_phase<C>(base, Y, SQR ? PH_SQR : PH_MUL);
    if (!MMA) {
        run_phase<C>(base, nullptr, PH_HIGH);
        run_phase<C>(base, nullptr, PH_LOW);
    } else {
        q1_to_bytes<C>(S, role, lane);
        phase_mma<C, true>(base);
        qhat_to_bytes<C>(S, role, lane);
        phase_mma<C, false>(base);
        low_to_value<C>(S, role, lane);
    }
}


// ---- witness step: exact (q, rem) of one mul_mod ------------------------------------------------------
// Replaces the witness computation of BigUintChip::mul_mod (called from /root/reference/src/paillier.rs:51,55,57
// through pow_mod_fixed_exp): q = floor(a b / n^2), rem = a b mod n^2, both canonical.
//
// Chain values are kept as strict digits of x * 2^s with 2s = sh_w, Nt_w = n^2 << sh_w (sh_w even).  Then
//   T = (a 2^s)(b 2^s) = (a b) 2^sh_w,   floor(T / Nt_w) = q,   T mod Nt_w = rem 2^sh_w,
// so the engine's own Barrett phases (A on IMAD, B and C on the tensor pipe) estimate q directly:
// q-hat is within one of q (fuzzed in tests/model_block28.py), V' = lo(T) - lo(q-hat Nt_w) = rem 2^sh_w + k Nt_w
// with k in {-1, 0}.  The tail below makes both exact, per lane, with all G warps working on their own block:
//   1. k estimated from the two top digits of V' (double precision, off by one only next to a multiple of Nt_w);
//   2. R = V' - k Nt_w and q = q-hat + k rippled to unsigned digits per block, block carries resolved through
//      shared flags (a carry that runs through a whole block costs one more round; rounds are counted with
//      __syncthreads_or, typically two);
//   3. range check R in [0, Nt_w) by per-block comparison; the rare miss adds/subtracts Nt_w once more;
//   4. q and rem = R >> sh_w packed into 64-bit words (the record), hashed, optionally stored;
//   5. the next chain operand, strict digits of rem 2^s = R >> s, written to `next_dst`.
template <class C> __device__ __forceinline__ int* west_ptr(Smem<C>& S) { return (int*)(S.rnt + (size_t)C::RTAB4 * 16); }

struct WStep {             // per-step outputs of one lane (null = not wanted)
    u64w* rec;             // 2*words_out words: q then rem
    u64w* rem_out;         // words_out words: rem only (the ciphertext of the final step)
    u64w* q_out = nullptr; // words_out words: q only
};

template <int L>
__device__ __forceinline__ u64w extract64(const int* F, int bit, int lane) {
    const int p = bit / W, off = bit - p * W;
    u64w w = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int pp = p + i;
        const u64w d = pp < L ? (u64w)(unsigned)F[pp * 32 + lane] : (u64w)0;
        const int sft = W * i - off;
        if (i == 0) w = d >> off;
        else if (sft < 64) w |= d << sft;
    }
    return w;
}

template <int BL>
__device__ __forceinline__ int add_carry_block(int (&d)[BL], int cin) {
    if (!cin) return 0;
    int c = cin;
#pragma unroll
    for (int k = 0; k < BL; k++) { int t = d[k] + c; d[k] = t & ((1 << W) - 1); c = t >> W; }
    return c;
}

template <class C>
__device__ __noinline__ u64w w_tail(int4* smem_base, int4* next_dst, WStep out, int sh, int words_out,
                                                  double inv, const u64w* __restrict__ cpow) {
    constexpr int G = C::G, BL = C::BL, L = C::L, MSK = (1 << W) - 1;
    Smem<C> S(smem_base);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* est = west_ptr<C>(S);
    int* scr = (int*)S.Q;                         // the s8 rows are dead once q-hat has been decoded
    int* cfR[2] = {scr, scr + G * 32};
    int* cfQ[2] = {scr + 2 * G * 32, scr + 3 * G * 32};
    int* cmpf = scr + 4 * G * 32;
    u64w* hsum = (u64w*)(scr + 5 * G * 32);
    const int* LO = lo_ptr<C>(S);
    const int* CA = ca_ptr<C>(S);
    const int* ntu = (const int*)S.two_sh + warp * C::CH * 4;      // unsigned digits of Nt_w, this block (broadcast reads)
    // raw digits of V' = lo(T) - lo(q-hat Nt_w) and the digits of q-hat, this block
    int rd[BL], qd[BL];
    {
        int a[C::CH * 4];
        load_block<C>(a, blk_ptr<C>(S.T, warp, lane), 32);
        const unsigned* row = (const unsigned*)(as_ptr<C>(S) + lane * C::RS) + warp * BL;
#pragma unroll
        for (int k = 0; k < BL; k++) {
            const int j = warp * BL + k;
            rd[k] = a[k] - LO[dl_index(j, lane)] - (j > 0 ? CA[dl_index(j, lane)] : 0);
            const unsigned v = row[k];
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
        // R -= adj * Nt_w, q += adj: local ripples to unsigned digits
        int cR = 0, cQ = (warp == 0) ? adj : 0;
        if (first || adj) {
#pragma unroll
            for (int k = 0; k < BL; k++) { int t = rd[k] - adj * ntu[k] + cR; rd[k] = t & MSK; cR = t >> W; }
#pragma unroll
            for (int k = 0; k < BL; k++) { int t = qd[k] + cQ; qd[k] = t & MSK; cQ = t >> W; }
        } else cQ = 0;
        first = false;
        for (;;) {             // carries between blocks (numbers are taken mod 2^(28L): the top block's carry leaves)
            if (warp == G - 1) { cR = 0; cQ = 0; }
            cfR[rnd][warp * 32 + lane] = cR; cfQ[rnd][warp * 32 + lane] = cQ;
            if (!__syncthreads_or((cR | cQ) != 0)) break;
            const int iR = warp ? cfR[rnd][(warp - 1) * 32 + lane] : 0, iQ = warp ? cfQ[rnd][(warp - 1) * 32 + lane] : 0;
            cR = add_carry_block<BL>(rd, iR);
            cQ = add_carry_block<BL>(qd, iQ);
            rnd ^= 1;
        }
        // R in [0, Nt_w)?  negative values show as a top digit >= 2^27
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
    // canonical digits, flat [digit][lane], over the T buffer: R then q
    int* Rf = (int*)S.T;
    int* Qf = Rf + L * 32;
#pragma unroll
    for (int k = 0; k < BL; k++) { Rf[(warp * BL + k) * 32 + lane] = rd[k]; Qf[(warp * BL + k) * 32 + lane] = qd[k]; }
    __syncthreads();
    // record words: q = bits [64j, 64j+64) of q, rem = the same bits of R >> sh
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

// one witnessed mul_mod: operands V (and Y, or V itself when sqr) are strict digits of a 2^s, b 2^s with a, b canonical;
// V is preserved.  Returns the record hash (valid in warp 0).
template <class C>
__device__ __forceinline__ u64w mulmod_w(int4* smem_base, const int4* Y, int sqr, int4* next_dst, WStep out,
                                                       int sh, int words_out, double inv, const u64w* cpow) {
    Smem<C> S(smem_base);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    phase_product<C>(smem_base, Y, sqr);
    q1_to_bytes<C>(S, role, lane);
    phase_mma<C, true>(smem_base);
    qhat_to_bytes<C>(S, role, lane);
    phase_mma<C, false>(smem_base);
    return w_tail<C>(smem_base, next_dst, out, sh, words_out, inv, cpow);
}

}  // namespace b28
}  // namespace pb200
