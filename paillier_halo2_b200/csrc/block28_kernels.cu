// block28_kernels.cu — kernels and host driver of the block28 engine (device arithmetic in block28.cuh).
//
// Replaces, for `count` independent units: g.modpow(m, n2), r.modpow(n, n2) and the final product of
// paillier_enc_native (/root/reference/src/paillier.rs:87-92), and the fold of paillier_add_native
// (:94-97).  Chain per unit (all modulo Nt, a multiple of n^2):
//   r^n : left-to-right sliding window (w = 6) over the per-key exponent n; the 32 odd powers of r live in
//         a per-CTA global scratch table (L2-resident), the window schedule is computed once per key;
//   g^m : fixed-base comb, w-bit windows (w = 12 for |n| >= 1024, else 8): product of TG[i][digit_i(m)],
//         TG[i][d] = g^(d * 2^(w i)) mod Nt, built once per key by this engine (k_gtable_*);
//   c   : gm * rn, then one exact canonicalisation (finalize) and the shift back from Nt to n^2.
#include "engine.hpp"
#include "block28.cuh"
#include "block28u.cuh"
#include <cstdio>
#include <vector>
#include <cstring>
#include <cstdlib>

namespace pb200 {
using namespace b28;

constexpr int WIN = 6;                 // sliding window of the r-chain (6: 293 + 31 multiplications at |n| = 2048; 5: 341 + 15)
constexpr int TABN = 1 << (WIN - 1);   // odd powers r^1, r^3, ..., r^(2*TABN-1)
constexpr int SCRATCH_ENTRIES = TABN + 2;   // + r^2 (table build) + rn (kept while g^m is computed)

struct B28Dev {            // per-key device-side descriptor (same for every configuration)
    const int4* consts;    // mu, Nt, two_sh : 3 * ENTRY4 int4
    const int4* uconsts2;  // the same for the two-lane-group variant
    const int4* uconsts;   // block28u: constants + CM tables (UL<C>::KEY_BYTES), null when the configuration has no tcgen05 variant
    const int2* ops;       // r-chain schedule: (number of squarings, table index or -1)
    int n_ops;
    int first_idx;         // table index the chain starts from
    const int4* tg;        // comb table: [window][2^comb_bits][ENTRY4]
    int n_windows;
    int comb_bits;         // window width of the fixed-base comb for g^m
    const int4* n_entry;   // g == n + 1 (the standard Paillier generator): strict digits of n, g^m = 1 + m n (mod n^2); else null
    int words_in, words_out;
    int sh;                // Nt = n2 << sh
    unsigned nt_top;       // floor(Nt / 2^(28(L-2)))
    unsigned sms;
};

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ unsigned smid() { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
template <class C>
__device__ __forceinline__ int& digit_ref(int4* buf, int p, int lane) {
    int blk = p / C::BL, k = p % C::BL;
    return ((int*)(buf + blk * C::BLK4 + (k >> 2) * 32 + lane))[k & 3];
}

// per-lane buffer <- strict digits of the unsigned integer in src[0..nwords) (u64 little-endian)
template <class C>
__device__ __forceinline__ void load_value(int4* buf, const u64* src, int nwords, int role, int lane) {
    int a[C::CH * 4];
    int carry = 0;
#pragma unroll
    for (int k = 0; k < C::BL; k++) {
        int bit = W * (role * C::BL + k);
        int wi = bit >> 6, sh = bit & 63;
        u64 lo = wi < nwords ? src[wi] : 0, hi = wi + 1 < nwords ? src[wi + 1] : 0;
        u64 v = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        int t = (int)(v & ((1u << W) - 1)) + carry;
        int d = sgxt28(t);
        carry = (t - d) >> W;
        a[k] = d;
    }
#pragma unroll
    for (int k = C::BL; k < C::CH * 4; k++) a[k] = 0;
    store_block<C>(blk_ptr<C>(buf, role, lane), a);
    __syncthreads();
    if (role + 1 < C::G) *(int*)blk_ptr<C>(buf, role + 1, lane) += carry;
    __syncthreads();
}

// The 32 inputs of a CTA step are contiguous in global memory: copy them with coalesced loads into a word-major staging area
// ([word][lane], row stride 33 to spread the transposing stores over the banks) and cut the digits out of shared memory
// (a lane reading its own 512-byte ciphertext word by word touches 32 different lines per instruction).
template <class C>
__device__ __forceinline__ void stage_words(u64* stage, const u64* src, int units_avail, int nwords) {
    const int total = 32 * nwords;
    for (int idx = threadIdx.x; idx < total; idx += C::THREADS) {
        const int l = idx / nwords, wi = idx - l * nwords;
        stage[wi * 33 + l] = l < units_avail ? src[(size_t)l * nwords + wi] : 0;
    }
    __syncthreads();
}
template <class C>
__device__ __forceinline__ void load_value_staged(int4* buf, const u64* stage, int nwords, int role, int lane) {
    int a[C::CH * 4];
    int carry = 0;
#pragma unroll
    for (int k = 0; k < C::BL; k++) {
        int bit = W * (role * C::BL + k);
        int wi = bit >> 6, sh = bit & 63;
        u64 lo = wi < nwords ? stage[wi * 33 + lane] : 0, hi = wi + 1 < nwords ? stage[(wi + 1) * 33 + lane] : 0;
        u64 v = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        int t = (int)(v & ((1u << W) - 1)) + carry;
        int d = sgxt28(t);
        carry = (t - d) >> W;
        a[k] = d;
    }
#pragma unroll
    for (int k = C::BL; k < C::CH * 4; k++) a[k] = 0;
    store_block<C>(blk_ptr<C>(buf, role, lane), a);
    __syncthreads();
    if (role + 1 < C::G) *(int*)blk_ptr<C>(buf, role + 1, lane) += carry;
    __syncthreads();
}

template <class C>
__device__ __forceinline__ void set_one(int4* buf, int role, int lane) {
    int a[C::CH * 4];
#pragma unroll
    for (int k = 0; k < C::CH * 4; k++) a[k] = 0;
    if (role == 0) a[0] = 1;
    store_block<C>(blk_ptr<C>(buf, role, lane), a);
    __syncthreads();
}

// smem per-lane value <-> global [entry][blk][chunk][lane] (coalesced)
template <class C>
__device__ __forceinline__ void copy_to_global(int4* g, const int4* buf, int role, int lane) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) g[(role * C::CH + c) * 32 + lane] = buf[(role * C::CH + c) * 32 + lane];
}
template <class C>
__device__ __forceinline__ void copy_from_global(int4* buf, const int4* g, int role, int lane) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) buf[(role * C::CH + c) * 32 + lane] = g[(role * C::CH + c) * 32 + lane];
    __syncthreads();
}
// smem per-lane value <- one table entry per lane ([blk][chunk] int4, gathered)
template <class C>
__device__ __forceinline__ void gather_entry(int4* buf, const int4* entry, int role, int lane) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) buf[(role * C::CH + c) * 32 + lane] = __ldg(entry + role * C::CH + c);
    __syncthreads();
}
template <class C>
__device__ __forceinline__ void scatter_entry(int4* entry, const int4* buf, int role, int lane) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) entry[role * C::CH + c] = buf[(role * C::CH + c) * 32 + lane];
}

template <class C>
__device__ __forceinline__ void load_consts(int4* smem_base, const B28Dev& K) {
    int4* k = smem_base + 5 * C::VAL4;
    for (int i = threadIdx.x; i < 3 * C::ENTRY4 + 2 * C::RTAB4; i += C::THREADS) k[i] = K.consts[i];
    __syncthreads();
}


// ---- engine selection: 0 = block28 (all phases on IMAD), 1 = block28t (mma.sync for phases B, C), 2 = block28u (tcgen05) ----------
template <class C, int ENG> struct View {
    typedef Smem<C> type;
    static constexpr size_t BYTES = C::SMEM_BYTES;
    static constexpr int PER_SM = C::CTAS_PER_SM;
    static constexpr int LG = 1, THREADS = C::THREADS;
};
template <class C> struct View<C, 2> {            // block28u, one lane group per CTA
    typedef SmemU<C, 1> type;
    static constexpr size_t BYTES = UL<C, 1>::SMEM_BYTES;
    static constexpr int PER_SM = UL<C, 1>::CTAS_PER_SM;
    static constexpr int LG = 1, THREADS = C::THREADS;
};
template <class C> struct View<C, 4> : View<C, 2> {};      // block28u as the tally runs it: MMAs issued by thread 0 (see phases_bc_umma)
template <class C> struct View<C, 3> {            // block28u, two lane groups (64 ciphertexts) per CTA
    typedef SmemU<C, 2> type;
    static constexpr size_t BYTES = UL<C, 2>::SMEM_BYTES;
    static constexpr int PER_SM = UL<C, 2>::CTAS_PER_SM;
    static constexpr int LG = 2, THREADS = 2 * C::THREADS;
};
template <class C, bool SQR, int ENG, class SV>
__device__ __forceinline__ void mm(SV& S, const int4* Y, int role, int lane) {
    if constexpr (ENG >= 2) mulmod_u<C, View<C, ENG>::LG, SQR, ENG != 4>(S, Y);
    else mulmod<C, SQR, ENG == 1>(S, Y, role, lane);
}
// per-key constants into shared memory (+ TMEM and mbarriers for block28u)
template <class C, int ENG, class SV>
__device__ __forceinline__ void cta_begin(SV& S, int4* smem_base, const B28Dev& K) {
    if constexpr (ENG >= 2) {
        typedef UL<C, View<C, ENG>::LG> U;
        int4* dst = (int4*)((unsigned char*)smem_base + U::OFF_CONST);
        const int4* src = ENG == 3 ? K.uconsts2 : K.uconsts;
        for (int i = threadIdx.x; i < U::KEY_BYTES / 16; i += U::THREADS) dst[i] = src[i];
        umma_setup<C, View<C, ENG>::LG>(S);
    } else load_consts<C>(smem_base, K);
}
template <class C, int ENG, class SV>
__device__ __forceinline__ void cta_end(SV& S) {
    if constexpr (ENG >= 2) umma_teardown<C, View<C, ENG>::LG>(S);
}

// V <- V * (constant in shared memory, ENTRY4 int4, broadcast)
template <class C, int ENG, class SV>
__device__ __forceinline__ void mulmod_const(SV& S, const int4* cst, int role, int lane) {
    // expand the constant into the per-lane B buffer (keeps phase_product on one code path)
#pragma unroll
    for (int c = 0; c < C::CH; c++) S.B[(role * C::CH + c) * 32 + lane] = cst[role * C::CH + c];
    __syncthreads();
    mm<C, false, ENG>(S, S.B, role, lane);
}

// Exact canonicalisation of the lazy value in V (one thread per lane, role 0): out = ((V * 2^sh) mod Nt) >> sh,
// written as words_out u64 words.  V must already have been multiplied by two_sh.
template <class C>
__device__ void canonical_out(int4* V, const int4* NtC, const B28Dev& K, u64* out, int lane) {
    constexpr int L = C::L;
    const int* nt = (const int*)NtC;
    auto NT = [&](int p) { return nt[(p / C::BL) * C::CH * 4 + (p % C::BL)]; };
    auto D = [&](int p) -> int& { return digit_ref<C>(V, p, lane); };
    // full ripple to strict centred digits
    int carry = 0;
    for (int p = 0; p < L; p++) { int t = D(p) + carry; int d = sgxt28(t); carry = (t - d) >> W; D(p) = d; }
    // quotient estimate from the two top digits
    long long vt = ((long long)D(L - 1) << W) + D(L - 2);
    long long q = vt / (long long)K.nt_top;
    if (vt < 0 && vt % (long long)K.nt_top) q -= 1;           // floor division
    // v -= q * Nt, then fix up into [0, Nt)
    long long c64 = 0;
    for (int p = 0; p < L; p++) {
        long long t = (long long)D(p) - q * (long long)NT(p) + c64;
        int d = sgxt28((int)t); c64 = (t - d) >> W; D(p) = d;
    }
    for (int it = 0; it < 6; it++) {
        int top = 0;
        for (int p = L - 1; p >= 0; p--) { top = D(p); if (top) break; }
        int dir;
        if (top < 0) dir = 1;                                  // negative: add Nt
        else {
            // v >= Nt ?  compare via sign of v - Nt
            int cc = 0, last = 0; bool nonzero = false;
            for (int p = 0; p < L; p++) { int t = D(p) - NT(p) + cc; int d = sgxt28(t); cc = (t - d) >> W; if (d) { last = d; nonzero = true; } }
            dir = (!nonzero || last > 0) ? -1 : 0;             // v - Nt >= 0: subtract
        }
        if (dir == 0) break;
        int cc = 0;
        for (int p = 0; p < L; p++) { int t = D(p) + dir * NT(p) + cc; int d = sgxt28(t); cc = (t - d) >> W; D(p) = d; }
    }
    // unsigned digits (value is now in [0, Nt))
    carry = 0;
    for (int p = 0; p < L; p++) { int t = D(p) + carry; D(p) = t & ((1 << W) - 1); carry = t >> W; }
    // shift right by sh and pack into 64-bit words
    for (int j = 0; j < K.words_out; j++) {
        u64 w = 0;
        int bit0 = 64 * j + K.sh;
        int p0 = bit0 / W;
        for (int p = p0; p < p0 + 4 && p < L; p++) {
            int off = W * p - bit0;                           // position of digit p relative to the word
            u64 d = (u64)(unsigned)D(p);
            if (off >= 64) break;
            w |= off >= 0 ? d << off : d >> (-off);
        }
        out[j] = w;
    }
}

template <class C, int ENG, class SV>
__device__ __forceinline__ void finalize(SV& S, const B28Dev& K, u64* out, bool active, int role, int lane) {
    mulmod_const<C, ENG>(S, S.two_sh, role, lane);
    if (role == 0 && active) canonical_out<C>(S.V, S.Nt, K, out, lane);
    __syncthreads();
}

// ---- kernels ---------------------------------------------------------------------------------
// comb table, stage 1: bases[i] = g^(2^(w i)) mod Nt, i < n_windows (one CTA; every lane computes the same value)
template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) k_gtable_bases(B28Dev K, const u64* g_words, int4* bases) {
    extern __shared__ int4 smem[];
    Smem<C> S(smem);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    load_consts<C>(smem, K);
    load_value<C>(S.V, g_words, K.words_in, role, lane);
    for (int i = 0; i < K.n_windows; i++) {
        if (lane == 0) scatter_entry<C>(bases + (size_t)i * C::ENTRY4, S.V, role, lane);
        __syncthreads();
        if (i + 1 < K.n_windows)
            for (int s = 0; s < K.comb_bits; s++) mulmod<C, true>(S, nullptr, role, lane);
    }
}
// comb table, stage 2: lane = window; TG[i][d] = bases[i]^d, d < 2^w
template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) k_gtable_fill(B28Dev K, const int4* bases, int4* tg) {
    extern __shared__ int4 smem[];
    Smem<C> S(smem);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    load_consts<C>(smem, K);
    int win = blockIdx.x * 32 + lane;
    bool active = win < K.n_windows;
    if (!active) win = K.n_windows - 1;
    const int nd = 1 << K.comb_bits;
    int4* row = tg + (size_t)win * nd * C::ENTRY4;
    set_one<C>(S.V, role, lane);
    if (active) scatter_entry<C>(row, S.V, role, lane);
    gather_entry<C>(S.V, bases + (size_t)win * C::ENTRY4, role, lane);
    gather_entry<C>(S.B, bases + (size_t)win * C::ENTRY4, role, lane);
    for (int d = 1; d < nd; d++) {
        if (active) scatter_entry<C>(row + (size_t)d * C::ENTRY4, S.V, role, lane);
        __syncthreads();
        if (d + 1 < nd) mulmod<C, false>(S, S.B, role, lane);
    }
}

// Per-CTA global scratch is sized by the CTAs that can be RESIDENT, not by the grid: a CTA takes one of the `per_sm` slots of the
// SM it runs on (a bit of masks[%smid], atomicCAS) and gives it back when it ends.  per_sm is the occupancy the runtime reports
// for the kernel, so a free bit always exists; the loop still re-reads the mask should that ever not hold.
struct SlotPool { unsigned* masks; int per_sm; };
__device__ __forceinline__ int slot_acquire(const SlotPool& P) {
    unsigned* mk = P.masks + smid();
    const unsigned full = P.per_sm >= 32 ? 0xffffffffu : ((1u << P.per_sm) - 1u);
    for (;;) {
        const unsigned cur = *(volatile unsigned*)mk;
        const unsigned free_bits = ~cur & full;
        if (!free_bits) { __nanosleep(100); continue; }
        const int b = __ffs((int)free_bits) - 1;
        if (atomicCAS(mk, cur, cur | (1u << b)) == cur) return (int)smid() * P.per_sm + b;
    }
}
__device__ __forceinline__ void slot_release(const SlotPool& P, int slot) {
    atomicAnd(P.masks + slot / P.per_sm, ~(1u << (slot % P.per_sm)));
}
__global__ void k_nsmid(unsigned* out) { unsigned v; asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v)); *out = v; }

template <class C, int ENG>
__global__ void __launch_bounds__(View<C, ENG>::THREADS, View<C, ENG>::PER_SM) k_encrypt(B28Dev K, const u64* __restrict__ m, const u64* __restrict__ r,
                                                            size_t count, u64* __restrict__ c_out, int4* scratch, SlotPool pool) {
    extern __shared__ int4 smem[];
    __shared__ int s_slot[2];
    constexpr int LG = View<C, ENG>::LG;
    typename View<C, ENG>::type S(smem);
    const int lane = threadIdx.x & 31, role = (threadIdx.x >> 5) % C::G, grp = (threadIdx.x >> 5) / C::G;      // grp: lane group (LG per CTA)
    if (threadIdx.x == 0) for (int g = 0; g < LG; g++) s_slot[g] = slot_acquire(pool);
    cta_begin<C, ENG>(S, smem, K);
    const int slot = s_slot[grp];
    size_t unit = (size_t)blockIdx.x * (32 * LG) + grp * 32 + lane;
    const bool active = unit < count;
    if (!active) unit = count - 1;
    int4* tab = scratch + (size_t)slot * SCRATCH_ENTRIES * C::VAL4;
    // ---- r^n: odd powers table, then the per-key window schedule
    load_value<C>(S.V, r + unit * K.words_in, K.words_in, role, lane);
    copy_to_global<C>(tab, S.V, role, lane);                                   // tab[0] = r
    mm<C, true, ENG>(S, nullptr, role, lane);                                   // r^2
    copy_to_global<C>(tab + (size_t)TABN * C::VAL4, S.V, role, lane);
    __syncthreads();
    copy_from_global<C>(S.V, tab, role, lane);
    for (int j = 1; j < TABN; j++) {
        copy_from_global<C>(S.B, tab + (size_t)TABN * C::VAL4, role, lane);
        mm<C, false, ENG>(S, S.B, role, lane);                                  // r^(2j+1)
        copy_to_global<C>(tab + (size_t)j * C::VAL4, S.V, role, lane);
    }
    __syncthreads();
    copy_from_global<C>(S.V, tab + (size_t)K.first_idx * C::VAL4, role, lane);
    for (int o = 0; o < K.n_ops; o++) {
        int2 op = K.ops[o];
        for (int s = 0; s < op.x; s++) mm<C, true, ENG>(S, nullptr, role, lane);
        if (op.y >= 0) {
            copy_from_global<C>(S.B, tab + (size_t)op.y * C::VAL4, role, lane);
            mm<C, false, ENG>(S, S.B, role, lane);
        }
    }
    copy_to_global<C>(tab + (size_t)(TABN + 1) * C::VAL4, S.V, role, lane);    // rn
    __syncthreads();
    // ---- g^m: comb over the bytes of m
    const u64* mw = m + unit * K.words_in;
    const int cb = K.comb_bits;
    auto comb_digit = [&](int i) -> int {
        const int bit = i * cb, w = bit >> 6, sft = bit & 63;
        u64 v = mw[w] >> sft;
        if (sft + cb > 64 && w + 1 < K.words_in) v |= mw[w + 1] << (64 - sft);
        return (int)(v & ((1u << cb) - 1));
    };
    if (K.n_entry) {
        // g = n + 1: (1 + n)^m = 1 + m n (mod n^2) — one multiplication by the per-key constant n instead of the comb
        load_value<C>(S.V, mw, K.words_in, role, lane);
        mulmod_const<C, ENG>(S, K.n_entry, role, lane);
        if (role == 0) ((int*)blk_ptr<C>(S.V, 0, lane))[0] += 1;
        __syncthreads();
    } else {
        gather_entry<C>(S.V, K.tg + (size_t)comb_digit(0) * C::ENTRY4, role, lane);
        for (int i = 1; i < K.n_windows; i++) {
            gather_entry<C>(S.B, K.tg + (((size_t)i << cb) + comb_digit(i)) * C::ENTRY4, role, lane);
            mm<C, false, ENG>(S, S.B, role, lane);
        }
    }
    // ---- c = gm * rn, canonical
    copy_from_global<C>(S.B, tab + (size_t)(TABN + 1) * C::VAL4, role, lane);
    mm<C, false, ENG>(S, S.B, role, lane);
    finalize<C, ENG>(S, K, c_out + unit * K.words_out, active, role, lane);
    if (threadIdx.x == 0) for (int g = 0; g < LG; g++) slot_release(pool, s_slot[g]);
    cta_end<C, ENG>(S);
}

// out = base^e mod n^2 for a per-key exponent e: the r-chain of k_encrypt (odd-power table in the CTA's scratch slot, sliding-window
// schedule computed once per key) with its own schedule.  Decryption's c^lambda mod n^2.
struct PowSched { const int2* ops; int n_ops; int first_idx; };
template <class C, int ENG>
__global__ void __launch_bounds__(View<C, ENG>::THREADS, View<C, ENG>::PER_SM) k_pow(B28Dev K, PowSched E, const u64* __restrict__ base, int base_words,
                                                                    size_t count, u64* __restrict__ out, int4* scratch, SlotPool pool) {
    extern __shared__ int4 smem[];
    __shared__ int s_slot[2];
    constexpr int LG = View<C, ENG>::LG;
    typename View<C, ENG>::type S(smem);
    const int lane = threadIdx.x & 31, role = (threadIdx.x >> 5) % C::G, grp = (threadIdx.x >> 5) / C::G;      // grp: lane group (LG per CTA)
    if (threadIdx.x == 0) for (int g = 0; g < LG; g++) s_slot[g] = slot_acquire(pool);
    cta_begin<C, ENG>(S, smem, K);
    const int slot = s_slot[grp];
    size_t unit = (size_t)blockIdx.x * (32 * LG) + grp * 32 + lane;
    const bool active = unit < count;
    if (!active) unit = count - 1;
    int4* tab = scratch + (size_t)slot * SCRATCH_ENTRIES * C::VAL4;
    load_value<C>(S.V, base + unit * base_words, base_words, role, lane);
    copy_to_global<C>(tab, S.V, role, lane);                                        // tab[0] = x
    mm<C, true, ENG>(S, nullptr, role, lane);                                   // x^2
    copy_to_global<C>(tab + (size_t)TABN * C::VAL4, S.V, role, lane);
    __syncthreads();
    copy_from_global<C>(S.V, tab, role, lane);
    for (int j = 1; j < TABN; j++) {
        copy_from_global<C>(S.B, tab + (size_t)TABN * C::VAL4, role, lane);
        mm<C, false, ENG>(S, S.B, role, lane);                                  // x^(2j+1)
        copy_to_global<C>(tab + (size_t)j * C::VAL4, S.V, role, lane);
    }
    __syncthreads();
    if (E.first_idx < 0) set_one<C>(S.V, role, lane);                               // e == 0
    else copy_from_global<C>(S.V, tab + (size_t)E.first_idx * C::VAL4, role, lane);
    for (int o = 0; o < E.n_ops; o++) {
        const int2 op = E.ops[o];
        for (int s_ = 0; s_ < op.x; s_++) mm<C, true, ENG>(S, nullptr, role, lane);
        if (op.y >= 0) {
            copy_from_global<C>(S.B, tab + (size_t)op.y * C::VAL4, role, lane);
            mm<C, false, ENG>(S, S.B, role, lane);
        }
    }
    finalize<C, ENG>(S, K, out + unit * K.words_out, active, role, lane);
    if (threadIdx.x == 0) for (int g = 0; g < LG; g++) slot_release(pool, s_slot[g]);
    cta_end<C, ENG>(S);
}

// ---- tally: the N-ary fold of paillier_add_native (/root/reference/src/paillier.rs:94-97) in ONE launch per GPU -------------------
// 1. every lane of every CTA folds a strided subset of the inputs (main loop, all lanes useful);
// 2. the CTAs fold pairwise up a binary tree by "last arriver continues": a CTA stores its 32 lane values in its node slot, bumps
//    the node counter and leaves if its sibling has not arrived yet; the one that arrives second loads the sibling's values and
//    multiplies (32 useful lanes per multiplication, no spinning, no co-residency assumption);
// 3. the one surviving CTA folds its 32 lanes (5 steps) and canonicalises;
// 4. multi-GPU (world > 1): it stores the canonical partial into every peer's mailbox over NVLink (plain stores to mapped peer
//    memory, then a release flag at system scope), waits for the peers' partials in its own mailbox, folds the `world` partials
//    (3 steps at 8 GPUs) and canonicalises again — every GPU ends with the full product, no host hop and no second launch.
constexpr int TALLY_MAXW = 8;                          // GPUs of one NVSwitch domain
constexpr int MAIL_WORDS = 160;                        // one partial as lazy digits: ENTRY4 int4 = up to 16 blocks x 5 chunks x 16 B = 1280 B
constexpr size_t MAIL_FLAG_OFF = (size_t)2 * TALLY_MAXW * MAIL_WORDS;      // u64 index of flags[parity][src]
constexpr size_t MAIL_U64 = MAIL_FLAG_OFF + 2 * TALLY_MAXW;
struct TallyPeer {
    int world, rank;
    unsigned long long epoch;          // sequence number of this collective call (>= 1); parity selects the mailbox half
    u64* mail[TALLY_MAXW];             // every rank's mailbox as mapped on THIS device (mail[rank] is the local one)
    int* flags;                        // the key's flag word (bit 2: a peer's partial did not arrive)
};

template <class C>
__device__ __forceinline__ void copy_from_global_cg(int4* buf, const int4* g, int role, int lane) {     // written by another SM: L2 only
#pragma unroll
    for (int c = 0; c < C::CH; c++) buf[(role * C::CH + c) * 32 + lane] = __ldcg(g + (role * C::CH + c) * 32 + lane);
    __syncthreads();
}
template <class C>
__device__ __forceinline__ void set_one_where(int4* buf, bool cond, int role, int lane) {
    if (cond) {
        int a[C::CH * 4];
#pragma unroll
        for (int k = 0; k < C::CH * 4; k++) a[k] = 0;
        if (role == 0) a[0] = 1;
        store_block<C>(blk_ptr<C>(buf, role, lane), a);
    }
    __syncthreads();
}
// V[lane] <- product over the lanes of its aligned group of `width` lanes (width a power of two <= 32)
template <class C, int ENG, class SV>
__device__ __forceinline__ void fold_lanes(SV& S, int width, int role, int lane) {
    for (int off = width >> 1; off >= 1; off >>= 1) {
#pragma unroll
        for (int ch = 0; ch < C::CH; ch++) S.B[(role * C::CH + ch) * 32 + lane] = S.V[(role * C::CH + ch) * 32 + (lane ^ off)];
        __syncthreads();
        mm<C, false, ENG>(S, S.B, role, lane);
    }
}

template <class C, int ENG>
__global__ void __launch_bounds__(C::THREADS, View<C, ENG>::PER_SM) k_tally(B28Dev K, const u64* __restrict__ c, size_t count, u64* __restrict__ out,
                                                                      int4* nodes, unsigned* node_cnt, TallyPeer P) {
    extern __shared__ int4 smem[];
    __shared__ int s_first;
    typename View<C, ENG>::type S(smem);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    cta_begin<C, ENG>(S, smem, K);
    // ---- 1. main loop
    const size_t stride = (size_t)gridDim.x * 32;
    const size_t first = (size_t)blockIdx.x * 32 + lane;
    const size_t max_iters = (count + stride - 1) / stride;
    set_one<C>(S.V, role, lane);
    for (size_t it = 0; it < max_iters; it++) {
        const size_t u = first + it * stride;
        const size_t base = (size_t)blockIdx.x * 32 + it * stride;                  // first unit of this CTA step
        stage_words<C>((u64*)S.T, c + base * K.words_out, base < count ? (int)(count - base < 32 ? count - base : 32) : 0, K.words_out);
        load_value_staged<C>(S.B, (const u64*)S.T, K.words_out, role, lane);
        set_one_where<C>(S.B, u >= count, role, lane);                               // lanes without an input multiply by one
        mm<C, false, ENG>(S, S.B, role, lane);
    }
    // ---- 2. tree over the CTAs
    {
        int idx = blockIdx.x, n = gridDim.x;
        size_t node_base = 0;
        while (n > 1) {
            const int j = idx >> 1;
            if ((idx ^ 1) < n) {
                int4* slot = nodes + (node_base + j) * 2 * (size_t)C::VAL4;
                copy_to_global<C>(slot + (size_t)(idx & 1) * C::VAL4, S.V, role, lane);
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) s_first = atomicAdd(node_cnt + node_base + j, 1u) == 0u;
                __syncthreads();
                if (s_first) { cta_end<C, ENG>(S); return; }                                              // the sibling will pick this value up
                if (threadIdx.x == 0) node_cnt[node_base + j] = 0;                   // both arrived: ready for the next launch
                __threadfence();
                copy_from_global_cg<C>(S.B, slot + (size_t)((idx & 1) ^ 1) * C::VAL4, role, lane);
                mm<C, false, ENG>(S, S.B, role, lane);
            }
            node_base += (size_t)((n + 1) >> 1);
            idx = j; n = (n + 1) >> 1;
        }
    }
    // ---- 3. the surviving CTA: fold its lanes, canonical partial
    fold_lanes<C, ENG>(S, 32, role, lane);
    if (P.world <= 1) { finalize<C, ENG>(S, K, out, lane == 0, role, lane); cta_end<C, ENG>(S); return; }
    // ---- 4. exchange over peer memory and combine.  What crosses NVLink is lane 0's LAZY value (strict digits, ENTRY4 int4): the
    // receivers multiply digit arrays directly, so neither a canonicalisation before the exchange nor a digit conversion after it
    const int par = (int)(P.epoch & 1);
    {
        const int4* v0 = S.V;                                   // lane 0's chunks: S.V[(block * CH + chunk) * 32 + 0]
        for (int idx = threadIdx.x; idx < P.world * C::ENTRY4; idx += C::THREADS) {
            const int peer = idx / C::ENTRY4, e = idx - peer * C::ENTRY4;
            int4* dst = reinterpret_cast<int4*>(P.mail[peer] + ((size_t)par * TALLY_MAXW + P.rank) * MAIL_WORDS) + e;
            *dst = v0[e * 32];
        }
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
        // publish: my partial is in peer t's mailbox;   then wait for peer t's partial in mine
        u64* theirs = P.mail[threadIdx.x] + MAIL_FLAG_OFF + (size_t)par * TALLY_MAXW + P.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(theirs), "l"(P.epoch) : "memory");
        const u64* flag = P.mail[P.rank] + MAIL_FLAG_OFF + (size_t)par * TALLY_MAXW + threadIdx.x;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= P.epoch) break;
            if (clock64() - t0 > 6000000000ll) { atomicOr(P.flags, 4); break; }      // ~3 s: a peer never called
            __nanosleep(200);
        }
    }
    __syncthreads();
    {
        const int src = lane < P.world ? lane : P.world - 1;
        const int4* e = reinterpret_cast<const int4*>(P.mail[P.rank] + ((size_t)par * TALLY_MAXW + src) * MAIL_WORDS);
#pragma unroll
        for (int c = 0; c < C::CH; c++) S.V[(role * C::CH + c) * 32 + lane] = __ldcg(e + role * C::CH + c);
        __syncthreads();
    }
    set_one_where<C>(S.V, lane >= P.world, role, lane);
    int width = 1; while (width < P.world) width <<= 1;
    fold_lanes<C, ENG>(S, width, role, lane);
    finalize<C, ENG>(S, K, out, lane == 0, role, lane);
    cta_end<C, ENG>(S);
}



// ---- diagnostic: one modular multiplication on raw lazy digits, on a chosen engine ---------------------------------------------
// v_in / y_in / v_out: one CTA's values in the shared-memory image layout ([block][chunk][lane] int4, VAL4 int4); t_out: the 2L-digit
// product after phase A (2 VAL4 int4); rows_out: the q-hat digits as packed s8 words, [lane][L].  Used by the parity tests to compare
// the engines digit for digit (the lazy digits of block28t and block28u are specified to be identical).
template <class C, int ENG>
__global__ void __launch_bounds__(View<C, ENG>::THREADS, View<C, ENG>::PER_SM) k_mulmod_dbg(B28Dev K, const int4* v_in, const int4* y_in, int reps,
                                                                                 int4* v_out, int4* t_out, unsigned* rows_out) {
    extern __shared__ int4 smem[];
    typename View<C, ENG>::type S(smem);
    constexpr int LG = View<C, ENG>::LG;
    const int lane = threadIdx.x & 31, role = (threadIdx.x >> 5) % C::G, grp = (threadIdx.x >> 5) / C::G;      // every lane group works on the same input
    cta_begin<C, ENG>(S, smem, K);
    copy_from_global<C>(S.V, v_in, role, lane);
    if (t_out) {            // phase A alone
        if (y_in) copy_from_global<C>(S.B, y_in, role, lane);
        if constexpr (ENG >= 2) phase_product_u<C, LG>(smem, S.B, y_in ? 0 : 1, S.tmem);
        else if constexpr (ENG == 1) phase_product<C>(S.V, S.B, y_in ? 0 : 1);
        else run_phase<C>(S.V, S.B, y_in ? PH_MUL : PH_SQR);
        if constexpr (ENG >= 2) {
            if (grp == LG - 1) for (int i = threadIdx.x % C::THREADS; i < 2 * C::VAL4; i += C::THREADS) t_out[i] = *(S.tblk(i / C::BLK4, 0) + i % C::BLK4);
        } else for (int i = threadIdx.x; i < 2 * C::VAL4; i += C::THREADS) t_out[i] = S.T[i];
        __syncthreads();
    } else {
        for (int r = 0; r < reps; r++) {
            if (y_in) { copy_from_global<C>(S.B, y_in, role, lane); mm<C, false, ENG>(S, S.B, role, lane); }
            else mm<C, true, ENG>(S, nullptr, role, lane);
        }
        if (grp == LG - 1) copy_to_global<C>(v_out, S.V, role, lane);
        if (rows_out && grp == LG - 1) {
#pragma unroll 1
            for (int k = 0; k < C::BL; k++) {
                const int qd = role * C::BL + k;
                unsigned w = 0;
                if constexpr (ENG >= 2) w = *(const unsigned*)(S.asc() + (UL<C, LG>::FRONT + (qd >> 2)) * UL<C, LG>::CHB + (grp * 32 + lane) * 16 + (qd & 3) * 4);
                else if constexpr (ENG == 1) w = ((const unsigned*)(as_ptr<C>(S) + lane * C::RS))[qd];
                rows_out[lane * C::L + qd] = w;
            }
        }
        __syncthreads();
    }
    cta_end<C, ENG>(S);
}


// diagnostic: `reps` modular squarings per CTA on a grid of CTAs, cycles of phase A and of phases B + C accumulated by thread 0
// (cyc[cta] = {phase A, phases B and C, whole loop}); every CTA works on the same input
template <class C, int ENG>
__global__ void __launch_bounds__(View<C, ENG>::THREADS, View<C, ENG>::PER_SM) k_mulmod_time(B28Dev K, const int4* v_in, int reps, int stagger, long long* cyc, unsigned* sm_count) {
    extern __shared__ int4 smem[];
    typename View<C, ENG>::type S(smem);
    constexpr int LG = View<C, ENG>::LG;
    const int lane = threadIdx.x & 31, role = (threadIdx.x >> 5) % C::G;
    cta_begin<C, ENG>(S, smem, K);
    copy_from_global<C>(S.V, v_in, role, lane);
    __shared__ unsigned s_order;
    if (threadIdx.x == 0) s_order = atomicAdd(sm_count + smid(), 1u);
    __syncthreads();
    if (stagger > 0 && (s_order & 1)) {             // delay the second CTA of every SM (diagnostic of the A / tensor ping-pong between co-resident CTAs)
        const long long t0 = clock64();
        while (clock64() - t0 < stagger) __nanosleep(200);
        __syncthreads();
    }
    long long ta = 0, tb = 0;
    const long long t_begin = clock64();
    // stagger -2: the first CTA of an SM runs phase A only, the second phases B + C only (interference between the two, no dependency);
    // -3: phase A only everywhere; -4: phases B + C only everywhere
    const bool do_a = stagger == -3 || (stagger == -2 && !(s_order & 1)) || stagger >= 0;
    const bool do_bc = stagger == -4 || (stagger == -2 && (s_order & 1)) || stagger >= 0;
    for (int r = 0; r < reps; r++) {
        const long long t0 = clock64();
        if (do_a) {
            if constexpr (ENG >= 2) phase_product_u<C, LG>(smem, nullptr, 1, S.tmem);
            else phase_product<C>(S.V, nullptr, 1);
        }
        const long long t1 = clock64();
        if (!do_bc) { ta += t1 - t0; continue; }
        if constexpr (ENG >= 2) phases_bc_umma<C, LG>(smem, S.tmem);
        else {
            q1_to_bytes<C>(S, role, lane);
            phase_mma<C, true>(S.V);
            qhat_to_bytes<C>(S, role, lane);
            phase_mma<C, false>(S.V);
            low_to_value<C>(S, role, lane);
        }
        const long long t2 = clock64();
        ta += t1 - t0; tb += t2 - t1;
    }
    if (threadIdx.x == 0) { cyc[3 * blockIdx.x] = ta; cyc[3 * blockIdx.x + 1] = tb; cyc[3 * blockIdx.x + 2] = clock64() - t_begin; }
    cta_end<C, ENG>(S);
}

// ---- witness kernels (reference chain with exact (q, rem) per mul_mod) ---------------------------------
struct WitDev {
    const int4* gtab;          // [n_bits][ENTRY4]: strict digits of (g^(2^i) mod n^2) * 2^s   (i = 0: g itself)
    const int4* one_s;         // ENTRY4: 2^s
    const u64* cpow;           // C^(j+1), j < 2*words_out (record hash, include/paillier_b200.h)
    const u64* n_words;        // exponent of the r-chain
    int exp_bits;              // bits(n)
    int sh;                    // Nt_w = n^2 << sh, sh even; chain values carry the factor 2^(sh/2)
    double inv;                // 2^(28(L-2)) / Nt_w
};
struct WitIO {
    const u64* m; const u64* r; size_t count;
    u64* c_out;                // nullable
    u64* records; const u64* offsets;   // nullable: unit u's records start at record index offsets[u]
    u64* digest;               // nullable
    int4* scratch;             // 2 values per CTA: acc of the r-chain, gm
};

// per-lane buffer <- strict digits of (src << shl)
template <class C>
__device__ __forceinline__ void load_value_shl(int4* buf, const u64* src, int nwords, int shl, int role, int lane) {
    int a[C::CH * 4];
    int carry = 0;
#pragma unroll
    for (int k = 0; k < C::BL; k++) {
        const int bit = W * (role * C::BL + k) - shl;
        u64 v = 0;
        if (bit >= 0) {
            const int wi = bit >> 6, sh = bit & 63;
            const u64 lo = wi < nwords ? src[wi] : 0, hi = wi + 1 < nwords ? src[wi + 1] : 0;
            v = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        } else if (bit > -W) {
            v = src[0] << (-bit);
        }
        const int t = (int)(v & ((1u << W) - 1)) + carry;
        const int d = sgxt28(t);
        carry = (t - d) >> W;
        a[k] = d;
    }
#pragma unroll
    for (int k = C::BL; k < C::CH * 4; k++) a[k] = 0;
    store_block<C>(blk_ptr<C>(buf, role, lane), a);
    __syncthreads();
    if (role + 1 < C::G) *(int*)blk_ptr<C>(buf, role + 1, lane) += carry;
    __syncthreads();
}

// ---- witness engine selection: 0 = block28t arithmetic (mma.sync phases), 1 = block28u arithmetic (tcgen05 phases) ---------------
template <class C, int WENG> struct WView {
    typedef Smem<C> type;
    static constexpr size_t BYTES = C::SMEM_W_BYTES;
    static constexpr int PER_SM = C::CTAS_PER_SM;
};
template <class C> struct WView<C, 1> {
    typedef SmemU<C, 1, true> type;
    static constexpr size_t BYTES = UL<C, 1, true>::SMEM_BYTES;
    static constexpr int PER_SM = UL<C, 1, true>::CTAS_PER_SM;
};
template <class C, int WENG, class SV>
__device__ __forceinline__ void w_begin(SV& S, int4* smem_base, const B28Dev& K) {
    if constexpr (WENG == 1) {
        typedef UL<C, 1, true> U;
        int4* dst = (int4*)((unsigned char*)smem_base + U::OFF_CONST);
        for (int i = threadIdx.x; i < U::KEY_BYTES / 16; i += C::THREADS) dst[i] = K.uconsts[i];
        umma_setup<C, 1, true>(S);
    } else load_consts<C>(smem_base, K);
}
template <class C, int WENG, class SV>
__device__ __forceinline__ void w_end(SV& S) {
    if constexpr (WENG == 1) umma_teardown<C, 1, true>(S);
}
// one witnessed mul_mod (contract of mulmod_w, block28.cuh)
template <class C, int WENG, class SV>
__device__ __forceinline__ u64 wmm(int4* smem_base, SV& S, const int4* Y, int sqr, int4* next_dst, WStep out, int sh, int words_out,
                                   double inv, const u64* cpow) {
    if constexpr (WENG == 1) return mulmod_wu<C>(smem_base, S.tmem, Y, sqr, next_dst, out, sh, words_out, inv, cpow);
    else return mulmod_w<C>(smem_base, Y, sqr, next_dst, out, sh, words_out, inv, cpow);
}

// table entries of the witness chain from 64-bit words: entry i = strict digits of (value_i << shl), one thread per entry.
// value_0 = g (words_in words); value_i = rem of g-chain record i-1 (words_out words at gchain + (i-1)*2*wo + wo)
template <class C>
__global__ void k_wtab(const u64* __restrict__ g_words, int words_in, const u64* __restrict__ gchain, int words_out,
                       int n_entries, int shl, int4* __restrict__ gtab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const u64* src = i == 0 ? g_words : gchain + (size_t)(i - 1) * 2 * words_out + words_out;
    const int nwords = i == 0 ? words_in : words_out;
    int* dst = (int*)(gtab + (size_t)i * C::ENTRY4);
    for (int k = 0; k < C::ENTRY4 * 4; k++) dst[k] = 0;
    int carry = 0;
    for (int p = 0; p < C::L; p++) {
        const int bit = W * p - shl;
        u64 v = 0;
        if (bit >= 0) {
            const int wi = bit >> 6, sh = bit & 63;
            const u64 lo = wi < nwords ? src[wi] : 0, hi = wi + 1 < nwords ? src[wi + 1] : 0;
            v = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        } else if (bit > -W) {
            v = src[0] << (-bit);
        }
        const int t = (int)(v & ((1u << W) - 1)) + carry;
        const int d = sgxt28(t);
        carry = (t - d) >> W;
        dst[(p / C::BL) * C::CH * 4 + (p % C::BL)] = d;
    }
}

// per-key g-chain on the witness engine: record i = (q, rem) of (g^(2^i))^2, i < n_bits, and the table entries
// gtab[i] = strict digits of g^(2^i) * 2^s.  One CTA; every lane computes the same value, lane 0 writes.
template <class C, int WENG>
__global__ void __launch_bounds__(C::THREADS, 1) k_gchain_w(B28Dev K, WitDev Wd, const u64* __restrict__ g_words, int n_bits,
                                                            u64* __restrict__ gchain, int4* __restrict__ gtab) {
    extern __shared__ int4 smem[];
    typename WView<C, WENG>::type S(smem);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    w_begin<C, WENG>(S, smem, K);
    load_value_shl<C>(S.V, g_words, K.words_in, Wd.sh >> 1, role, lane);
    for (int i = 0; i < n_bits; i++) {
        if (lane == 0) scatter_entry<C>(gtab + (size_t)i * C::ENTRY4, S.V, role, lane);
        WStep o; o.rec = lane == 0 ? gchain + (size_t)i * 2 * K.words_out : nullptr; o.rem_out = nullptr;
        wmm<C, WENG>(smem, S, nullptr, 1, S.V, o, Wd.sh, K.words_out, Wd.inv, Wd.cpow);
    }
    w_end<C, WENG>(S);
}

template <class C>
__device__ __forceinline__ void bcast_entry(int4* buf, const int4* entry, int role, int lane) {
#pragma unroll
    for (int c = 0; c < C::CH; c++) buf[(role * C::CH + c) * 32 + lane] = __ldg(entry + role * C::CH + c);
    __syncthreads();
}

// One unit per lane, the reference's chain (src/paillier.rs:51,55,57 -> pow_mod_fixed_exp):
//   g-chain: the popcount(m) multiplications acc *= g^(2^i) (the squarings are per key: gtab);
//   r-chain: for every bit i of n, sqr_i = cur^2 and, if the bit is set, acc *= cur; then gm * rn.
// Lanes walk their own set bits of m (inactive lanes multiply by one and emit nothing); the r-chain is uniform.
// In the r-chain the multiplication of an iteration is computed BEFORE its squaring (cur stays in V), the
// records keep the reference's order (sqr_i, then mul_i).
template <class C, int WENG>
__global__ void __launch_bounds__(C::THREADS, WView<C, WENG>::PER_SM) k_witness(B28Dev K, WitDev Wd, WitIO A) {
    extern __shared__ int4 smem[];
    typename WView<C, WENG>::type S(smem);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    w_begin<C, WENG>(S, smem, K);
    size_t unit = (size_t)blockIdx.x * 32 + lane;
    const bool active = unit < A.count;
    if (!active) unit = A.count - 1;
    const int wo = K.words_out;
    const size_t recw = 2 * (size_t)wo;
    const u64* mw = A.m + unit * K.words_in;
    int pc = 0;
    for (int w = 0; w < K.words_in; w++) pc += __popcll(mw[w]);
    int maxpc = pc;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) maxpc = max(maxpc, __shfl_xor_sync(0xffffffffu, maxpc, off));
    u64* rec = (A.records && active) ? A.records + A.offsets[unit] * recw : nullptr;
    u64 D = PB200_DIGEST_INIT;                                    // running digest of the unit (kept by warp 0)
    int4* sc_acc = A.scratch + (size_t)blockIdx.x * 2 * C::VAL4;
    int4* sc_gm = sc_acc + C::VAL4;
    // ---- g-chain
    bcast_entry<C>(S.V, Wd.one_s, role, lane);
    {
        int wi = 0; u64 cur = mw[0];
        for (int t = 0; t < maxpc; t++) {
            const bool act = t < pc;
            const int4* e = Wd.one_s;
            if (act) {
                while (!cur) cur = mw[++wi];
                const int b = __ffsll((long long)cur) - 1;
                cur &= cur - 1;
                e = Wd.gtab + (size_t)(64 * wi + b) * C::ENTRY4;
            }
            bcast_entry<C>(S.B, e, role, lane);
            WStep o; o.rec = (act && rec) ? rec + (size_t)t * recw : nullptr; o.rem_out = nullptr;
            const u64 H = wmm<C, WENG>(smem, S, S.B, 0, S.V, o, Wd.sh, wo, Wd.inv, Wd.cpow);
            if (act) D = (D ^ H) * PB200_DIGEST_PRIME;
        }
    }
    copy_to_global<C>(sc_gm, S.V, role, lane);
    // ---- r-chain
    bcast_entry<C>(S.B, Wd.one_s, role, lane);
    copy_to_global<C>(sc_acc, S.B, role, lane);
    load_value_shl<C>(S.V, A.r + unit * K.words_in, K.words_in, Wd.sh >> 1, role, lane);
    size_t idx = (size_t)pc;
    for (int i = 0; i < Wd.exp_bits; i++) {
        const bool bit = (Wd.n_words[i >> 6] >> (i & 63)) & 1;
        u64 Hm = 0;
        if (bit) {
            copy_from_global<C>(S.B, sc_acc, role, lane);
            WStep o; o.rec = rec ? rec + (idx + 1) * recw : nullptr; o.rem_out = nullptr;
            Hm = wmm<C, WENG>(smem, S, S.B, 0, S.B, o, Wd.sh, wo, Wd.inv, Wd.cpow);
            copy_to_global<C>(sc_acc, S.B, role, lane);
        }
        WStep o; o.rec = rec ? rec + idx * recw : nullptr; o.rem_out = nullptr;
        const u64 Hs = wmm<C, WENG>(smem, S, nullptr, 1, S.V, o, Wd.sh, wo, Wd.inv, Wd.cpow);
        D = (D ^ Hs) * PB200_DIGEST_PRIME;
        if (bit) D = (D ^ Hm) * PB200_DIGEST_PRIME;
        idx += bit ? 2 : 1;
    }
    // ---- final: mul_mod(gm, rn)
    __syncthreads();
    copy_from_global<C>(S.V, sc_gm, role, lane);
    copy_from_global<C>(S.B, sc_acc, role, lane);
    {
        WStep o; o.rec = rec ? rec + idx * recw : nullptr;
        o.rem_out = (A.c_out && active) ? A.c_out + unit * wo : nullptr;
        const u64 H = wmm<C, WENG>(smem, S, S.B, 0, nullptr, o, Wd.sh, wo, Wd.inv, Wd.cpow);
        D = (D ^ H) * PB200_DIGEST_PRIME;
    }
    if (role == 0 && active && A.digest) A.digest[unit] = D;
    w_end<C, WENG>(S);
}

// paillier_add_native / PaillierChip::add (src/paillier.rs:62-85, :94-97) for `count` pairs on the witness engine:
// one exact mul_mod per lane, rem (and optionally q) as 64-bit words.  Inputs are c_words words each (zero-extended,
// src/paillier.rs:79-80) and need not be reduced; a quotient that does not fit words_out words raises the range flag.
template <class C, int WENG>
__global__ void __launch_bounds__(C::THREADS, WView<C, WENG>::PER_SM) k_add_w(B28Dev K, WitDev Wd, const u64* __restrict__ c1,
                                                                      const u64* __restrict__ c2, int c_words, size_t count,
                                                                      u64* __restrict__ out, u64* __restrict__ q_out, int* flags) {
    extern __shared__ int4 smem[];
    typename WView<C, WENG>::type S(smem);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    w_begin<C, WENG>(S, smem, K);
    const int wo = K.words_out;
    for (size_t first = (size_t)blockIdx.x * 32; first < count; first += (size_t)gridDim.x * 32) {
        size_t unit = first + lane;
        const bool active = unit < count;
        if (!active) unit = count - 1;
        load_value_shl<C>(S.V, c1 + unit * c_words, c_words, Wd.sh >> 1, role, lane);
        load_value_shl<C>(S.B, c2 + unit * c_words, c_words, Wd.sh >> 1, role, lane);
        // the record goes through a per-lane staging area only when q is wanted; rem alone is written directly
        WStep o; o.rec = nullptr; o.rem_out = active ? out + unit * wo : nullptr;
        o.q_out = (active && q_out) ? q_out + unit * wo : nullptr;
        wmm<C, WENG>(smem, S, S.B, 0, nullptr, o, Wd.sh, wo, Wd.inv, Wd.cpow);
        // q must fit words_out words (range check of assign_integer(q, 2*enc_bits)): canonical q digits are in T
        {
            const int* Qf = (const int*)S.T + C::L * 32;
            const int qbits = 64 * wo;
            int bad = 0;
#pragma unroll
            for (int k = 0; k < C::BL; k++) {
                const int p = role * C::BL + k;
                const unsigned d = (unsigned)Qf[p * 32 + lane];
                const int lo = W * p;
                if (lo >= qbits) bad |= d != 0;
                else if (lo + W > qbits) bad |= (d >> (qbits - lo)) != 0;
            }
            if (bad && active) atomicOr(flags, 1);
        }
        __syncthreads();
    }
    w_end<C, WENG>(S);
}

// ---- host side ---------------------------------------------------------------------------------
struct Block28Key {
    int G = 0, BL = 0;
    std::string name;
    B28Dev dev{};
    int4* d_consts = nullptr; int2* d_ops = nullptr; int4* d_tg = nullptr; u64* d_gwords = nullptr; int4* d_nentry = nullptr;
    int4* d_scratch = nullptr; size_t scratch_slots = 0;     // r-power tables, one per resident-CTA slot (SlotPool)
    unsigned* d_slot_masks = nullptr; int nsmid = 0, enc_per_sm = 0;
    int4* d_nodes = nullptr; unsigned* d_node_cnt = nullptr; size_t node_ctas = 0;   // tally tree: 2 values per node, arrival counters
    u64* d_mail = nullptr;                                   // tally mailbox (peer-visible), MAIL_U64 words
    TallyPeer peer{};                                        // world <= 1 until block28_tally_peer_connect
    void* ipc_opened[TALLY_MAXW] = {};                       // peer mailboxes opened with cudaIpcOpenMemHandle
    int device = 0;
    int2* d_pow_ops = nullptr; PowSched pow{};               // decryption exponent (block28_pow_prepare)
    bool pow_ready = false;
    int sms = 148;
    uint64_t n_sqr = 0, n_mul = 0;   // modular squarings / multiplications per encryption
    int eng = 1;                     // 0: all phases on IMAD (block28), 1: phases B, C on mma.sync (block28t), 2: on tcgen05, 32 ciphertexts per
                                     // CTA (block28u), 3: on tcgen05, 64 per CTA (block28u2; the tally and keys without that variant run as 2)
    bool has_u = false;              // a block28u variant is compiled for this configuration
    int4* d_uconsts = nullptr; int4* d_uconsts2 = nullptr; bool has_u2 = false;
    // witness engine (lazy: block28_witness_prepare)
    BigInt n; uint32_t n_bits = 0;
    bool wit_ready = false;
    int4* d_wuconsts = nullptr; bool has_wu = false;     // witness engine on the tcgen05 phases
    int4* d_wconsts = nullptr; int4* d_gtab = nullptr; int4* d_one_s = nullptr; u64* d_cpow = nullptr; u64* d_nwords = nullptr;
    int4* d_wscratch = nullptr; size_t wscratch_ctas = 0;
    B28Dev wdev{}; WitDev wit{};
};

template <class C>
static void to_entry(const BigInt& v, std::vector<int>& out) {   // centred digits in [blk][chunk*4] layout
    out.assign(C::ENTRY4 * 4, 0);
    int carry = 0;
    for (int p = 0; p < C::L; p++) {
        int t = (int)v.bits_at((size_t)W * p, W) + carry;
        int d = ((t + (1 << (W - 1))) & ((1 << W) - 1)) - (1 << (W - 1));
        carry = (t - d) >> W;
        out[(p / C::BL) * C::CH * 4 + (p % C::BL)] = d;
    }
    if (carry != 0 || v.bits() > (size_t)W * C::L - 1) throw std::runtime_error("block28: constant does not fit");
}

// signed 7-bit digits of a constant (4 per 28-bit digit, the top one absorbs the remainder)
template <class C>
static void to_k7(const std::vector<int>& entry, std::vector<signed char>& k7) {
    k7.assign(C::K7, 0);
    for (int p = 0; p < C::L; p++) {
        int d = entry[(p / C::BL) * C::CH * 4 + (p % C::BL)];
        for (int i = 0; i < 3; i++) { int e = ((d + 64) & 127) - 64; k7[4 * p + i] = (signed char)e; d = (d - e) >> 7; }
        if (d < -128 || d > 127) throw std::runtime_error("block28: constant digit does not split into s8");
        k7[4 * p + 3] = (signed char)d;
    }
}
// reversed, zero padded, byte-shifted s8 table of a constant's 7-bit digits (B operand of the IMMA phases)
template <class C>
static void to_rtab(const std::vector<int>& entry, std::vector<int>& out_words) {
    std::vector<signed char> k7;
    to_k7<C>(entry, k7);
    std::vector<signed char> rfull(C::XLEN + 8, 0);
    for (int j = 0; j < C::K7; j++) rfull[C::PAD7 + (C::K7 - 1 - j)] = k7[j];
    std::vector<signed char> tab((size_t)C::RTAB4 * 16, 0);
    for (int sft = 0; sft < 4; sft++)
        for (int x = 0; x < C::XLEN; x++) tab[(size_t)sft * C::XSTR + x] = x + sft < C::XLEN ? rfull[x + sft] : 0;
    out_words.resize((size_t)C::RTAB4 * 4);
    memcpy(out_words.data(), tab.data(), tab.size());
}

// left-to-right sliding window (width WIN) over a fixed exponent: returns the table index the chain starts from (-1 for e == 0)
// and the list of (squarings, table index to multiply by or -1)
static int window_schedule(const BigInt& e, std::vector<int2>& ops) {
    int first_idx = -1;
    int i = (int)e.bits() - 1, pending = 0; bool first = true;
    while (i >= 0) {
        if (!e.bit(i)) { pending++; i--; continue; }
        int j = i - WIN + 1; if (j < 0) j = 0;
        while (!e.bit(j)) j++;
        int val = 0; for (int b = i; b >= j; b--) val = (val << 1) | (e.bit(b) ? 1 : 0);
        int len = i - j + 1;
        if (first) { first_idx = (val - 1) / 2; first = false; }
        else ops.push_back(make_int2(pending + len, (val - 1) / 2));
        pending = 0; i = j - 1;
    }
    if (pending) ops.push_back(make_int2(pending, -1));
    return first_idx;
}

#define CUK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { *cuda_err = e_; block28_destroy(key); return nullptr; } } while (0)

template <class C>
static Block28Key* create_cfg(const BigInt& n, const BigInt& g, uint32_t n_bits, int device, cudaStream_t st, cudaError_t* cuda_err) {
    Block28Key* key = new Block28Key();
    key->G = C::G; key->BL = C::BL;
    key->n = n; key->n_bits = n_bits; key->device = device;
    key->name = "block28t<" + std::to_string(C::G) + "," + std::to_string(C::BL) + ">";      // block28_set_engine(-1) below picks the fastest
    cudaDeviceProp prop;
    CUK(cudaGetDeviceProperties(&prop, device));
    key->sms = prop.multiProcessorCount;
    BigInt n2 = BigInt::mul(n, n);
    int sh = C::KN - (int)n2.bits();
    BigInt Nt = BigInt::shl(n2, sh);
    BigInt mu = BigInt::div(BigInt::pow2(2 * (size_t)C::BETA), Nt);
    BigInt two_sh = BigInt::pow2(sh);
    std::vector<int> e_mu, e_nt, e_sh, all;
    to_entry<C>(mu, e_mu); to_entry<C>(Nt, e_nt); to_entry<C>(two_sh, e_sh);
    all.insert(all.end(), e_mu.begin(), e_mu.end());
    all.insert(all.end(), e_nt.begin(), e_nt.end());
    all.insert(all.end(), e_sh.begin(), e_sh.end());
    std::vector<int> r_mu, r_nt;
    to_rtab<C>(e_mu, r_mu); to_rtab<C>(e_nt, r_nt);
    all.insert(all.end(), r_mu.begin(), r_mu.end());
    all.insert(all.end(), r_nt.begin(), r_nt.end());
    CUK(cudaMalloc(&key->d_consts, all.size() * sizeof(int)));
    CUK(cudaMemcpyAsync(key->d_consts, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    // block28u: the same constants followed by the two Toeplitz core-matrix tables (image of shared memory from OFF_CONST on), for
    // one and for two lane groups per CTA
    auto u_image = [&](auto LGC, int4** d_out) -> cudaError_t {
        constexpr int LG = decltype(LGC)::value;
        typedef UL<C, LG> U;
        std::vector<signed char> img(U::KEY_BYTES, 0);
        memcpy(img.data(), all.data(), 3 * C::ENTRY4 * 16);
        std::vector<signed char> k7(C::K7);
        to_k7<C>(e_mu, k7);
        umma_cm_table<C, LG>(k7.data(), true, img.data() + (U::OFF_CMH - U::OFF_CONST));
        to_k7<C>(e_nt, k7);
        umma_cm_table<C, LG>(k7.data(), false, img.data() + (U::OFF_CML - U::OFF_CONST));
        cudaError_t e = cudaMalloc(d_out, img.size());
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(*d_out, img.data(), img.size(), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(st);
    };
    if constexpr (UL<C, 1>::SUPPORTED) {
        CUK(u_image(std::integral_constant<int, 1>{}, &key->d_uconsts));
        key->has_u = true;
        CUK((cudaFuncSetAttribute(k_encrypt<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UL<C, 1>::SMEM_BYTES)));
        CUK((cudaFuncSetAttribute(k_tally<C, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UL<C, 1>::SMEM_BYTES)));
        CUK((cudaFuncSetAttribute(k_pow<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UL<C, 1>::SMEM_BYTES)));
    }
    if constexpr (UL<C, 1>::SUPPORTED && UL<C, 2>::SUPPORTED) {
        CUK(u_image(std::integral_constant<int, 2>{}, &key->d_uconsts2));
        key->has_u2 = true;
        CUK((cudaFuncSetAttribute(k_encrypt<C, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UL<C, 2>::SMEM_BYTES)));
        CUK((cudaFuncSetAttribute(k_pow<C, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UL<C, 2>::SMEM_BYTES)));
    }
    block28_set_engine(key, -1);
    // sliding-window schedule for the exponent n
    std::vector<int2> ops;
    int first_idx = window_schedule(n, ops);
    int comb_bits = n_bits >= 1024 ? 12 : 8;
    if (const char* e = getenv("PB200_COMB_BITS")) { int v = atoi(e); if (v >= 4 && v <= 16) comb_bits = v; }
    const int n_windows = (int)((n_bits + comb_bits - 1) / comb_bits);
    const bool g_std = g == BigInt::add(n, BigInt(1)) && !getenv("PB200_NO_GSTD");
    key->n_sqr = 1; key->n_mul = (TABN - 1) + (g_std ? 1 : (uint64_t)(n_windows - 1)) + 1;   // r^2; table; comb (or m*n); gm*rn
    for (auto& o : ops) { key->n_sqr += (uint64_t)o.x; if (o.y >= 0) key->n_mul += 1; }
    CUK(cudaMalloc(&key->d_ops, (ops.size() + 1) * sizeof(int2)));
    if (!ops.empty()) CUK(cudaMemcpyAsync(key->d_ops, ops.data(), ops.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    uint32_t win = (n_bits + 63) / 64;
    std::vector<u64> gw(win);
    g.to_u64_le(gw.data(), win);
    CUK(cudaMalloc(&key->d_gwords, win * sizeof(u64)));
    CUK(cudaMemcpyAsync(key->d_gwords, gw.data(), win * sizeof(u64), cudaMemcpyHostToDevice, st));
    if (!g_std) CUK(cudaMalloc(&key->d_tg, ((size_t)n_windows << comb_bits) * C::ENTRY4 * sizeof(int4)));
    if (g_std) {
        std::vector<int> e_n;
        to_entry<C>(n, e_n);
        CUK(cudaMalloc(&key->d_nentry, e_n.size() * sizeof(int)));
        CUK(cudaMemcpyAsync(key->d_nentry, e_n.data(), e_n.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CUK(cudaStreamSynchronize(st));
    }
    B28Dev& K = key->dev;
    K.n_entry = key->d_nentry;
    K.consts = key->d_consts; K.uconsts = key->d_uconsts; K.uconsts2 = key->d_uconsts2; K.ops = key->d_ops; K.n_ops = (int)ops.size(); K.first_idx = first_idx;
    K.tg = key->d_tg; K.n_windows = n_windows; K.comb_bits = comb_bits; K.words_in = (int)win; K.words_out = (int)((2 * n_bits + 63) / 64);
    K.sh = sh;
    K.sms = (unsigned)key->sms;
    K.nt_top = (unsigned)BigInt::shr(Nt, (size_t)W * (C::L - 2)).bits_at(0, 32);
    // comb table
    CUK(cudaFuncSetAttribute(k_gtable_bases<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    CUK(cudaFuncSetAttribute(k_gtable_fill<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    CUK((cudaFuncSetAttribute(k_encrypt<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES)));
    CUK((cudaFuncSetAttribute(k_encrypt<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES)));
    CUK((cudaFuncSetAttribute(k_tally<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES)));
    CUK((cudaFuncSetAttribute(k_tally<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES)));
    int4* d_bases = nullptr;
    if (!g_std) {      // the comb table is only needed for a general g
        CUK(cudaMalloc(&d_bases, (size_t)n_windows * C::ENTRY4 * sizeof(int4)));
        k_gtable_bases<C><<<1, C::THREADS, C::SMEM_BYTES, st>>>(K, key->d_gwords, d_bases); count_launch();
        k_gtable_fill<C><<<(n_windows + 31) / 32, C::THREADS, C::SMEM_BYTES, st>>>(K, d_bases, key->d_tg); count_launch();
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (d_bases) cudaFree(d_bases);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { *cuda_err = e; block28_destroy(key); return nullptr; }
    return key;
}

#define CUW(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

template <class C>
static cudaError_t ensure_slots(Block28Key* key, cudaStream_t st) {
    if (!key->d_scratch) {
        // slots = SM ids x resident CTAs per SM (occupancy of the kernel as compiled), independent of the batch size
        unsigned* d_n = nullptr; unsigned h_n = 0;
        CUW(cudaMalloc(&d_n, sizeof(unsigned)));
        k_nsmid<<<1, 1, 0, st>>>(d_n); count_launch();
        CUW(cudaMemcpyAsync(&h_n, d_n, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CUW(cudaStreamSynchronize(st));
        cudaFree(d_n);
        int per_sm = 0;
        CUW((cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encrypt<C, 1>, C::THREADS, C::SMEM_BYTES)));
        if constexpr (UL<C, 1>::SUPPORTED) {
            int per_u = 0;
            CUW((cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_u, k_encrypt<C, 2>, C::THREADS, UL<C, 1>::SMEM_BYTES)));
            if (per_u > per_sm) per_sm = per_u;
            if (per_sm < 2) per_sm = 2;       // a CTA of two lane groups takes two slots
        }
        if (per_sm < 1 || per_sm > 32 || h_n == 0) return cudaErrorLaunchOutOfResources;
        key->nsmid = (int)h_n; key->enc_per_sm = per_sm;
        CUW(cudaMalloc(&key->d_slot_masks, h_n * sizeof(unsigned)));
        CUW(cudaMemsetAsync(key->d_slot_masks, 0, h_n * sizeof(unsigned), st));
        const size_t slots = (size_t)h_n * per_sm;
        CUW(cudaMalloc(&key->d_scratch, slots * SCRATCH_ENTRIES * C::VAL4 * sizeof(int4)));
        key->scratch_slots = slots;
    }
    return cudaSuccess;
}

template <class C>
static cudaError_t encrypt_cfg(Block28Key* key, const u64* d_m, const u64* d_r, size_t count, u64* d_c, cudaStream_t st) {
    CUW(ensure_slots<C>(key, st));
    const SlotPool pool{key->d_slot_masks, key->enc_per_sm};
    const size_t ctas = (count + 31) / 32;
    bool done = false;
    if constexpr (UL<C, 1>::SUPPORTED && UL<C, 2>::SUPPORTED) if (key->eng == 3 && key->has_u2) {
        k_encrypt<C, 3><<<(unsigned)((count + 63) / 64), 2 * C::THREADS, UL<C, 2>::SMEM_BYTES, st>>>(key->dev, d_m, d_r, count, d_c, key->d_scratch, pool);
        done = true;
    }
    if constexpr (UL<C, 1>::SUPPORTED) if (!done && key->eng >= 2 && key->has_u) {
        k_encrypt<C, 2><<<(unsigned)ctas, C::THREADS, UL<C, 1>::SMEM_BYTES, st>>>(key->dev, d_m, d_r, count, d_c, key->d_scratch, pool);
        done = true;
    }
    if (done) {}
    else if (key->eng >= 1) k_encrypt<C, 1><<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, st>>>(key->dev, d_m, d_r, count, d_c, key->d_scratch, pool);
    else k_encrypt<C, 0><<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, st>>>(key->dev, d_m, d_r, count, d_c, key->d_scratch, pool);
    count_launch();
    return cudaGetLastError();
}

template <class C>
static cudaError_t pow_cfg(Block28Key* key, const u64* d_base, int base_words, size_t count, u64* d_out, cudaStream_t st) {
    CUW(ensure_slots<C>(key, st));
    CUW((cudaFuncSetAttribute(k_pow<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES)));
    CUW((cudaFuncSetAttribute(k_pow<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES)));
    const SlotPool pool{key->d_slot_masks, key->enc_per_sm};
    const size_t ctas = (count + 31) / 32;
    bool done = false;
    if constexpr (UL<C, 1>::SUPPORTED && UL<C, 2>::SUPPORTED) if (key->eng == 3 && key->has_u2) {
        k_pow<C, 3><<<(unsigned)((count + 63) / 64), 2 * C::THREADS, UL<C, 2>::SMEM_BYTES, st>>>(key->dev, key->pow, d_base, base_words, count, d_out, key->d_scratch, pool);
        done = true;
    }
    if constexpr (UL<C, 1>::SUPPORTED) if (!done && key->eng >= 2 && key->has_u) {
        k_pow<C, 2><<<(unsigned)ctas, C::THREADS, UL<C, 1>::SMEM_BYTES, st>>>(key->dev, key->pow, d_base, base_words, count, d_out, key->d_scratch, pool);
        done = true;
    }
    if (done) {}
    else if (key->eng >= 1) k_pow<C, 1><<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, st>>>(key->dev, key->pow, d_base, base_words, count, d_out, key->d_scratch, pool);
    else k_pow<C, 0><<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, st>>>(key->dev, key->pow, d_base, base_words, count, d_out, key->d_scratch, pool);
    count_launch();
    return cudaGetLastError();
}

// nodes of the CTA tree over `n` leaves: sum of ceil(n / 2^l) for l >= 1
static size_t tree_nodes(size_t n) { size_t t = 0; while (n > 1) { n = (n + 1) >> 1; t += n; } return t; }

template <class C>
static cudaError_t tally_cfg(Block28Key* key, const u64* d_c, size_t count, u64* d_out, bool collective, cudaStream_t st) {
    const size_t cap = (size_t)C::CTAS_PER_SM * key->sms;
    size_t ctas = (count + 31) / 32;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    if (!key->d_nodes) {
        const size_t nodes = tree_nodes(cap) + 1;
        CUW(cudaMalloc(&key->d_nodes, nodes * 2 * C::VAL4 * sizeof(int4)));
        CUW(cudaMalloc(&key->d_node_cnt, nodes * sizeof(unsigned)));
        CUW(cudaMemsetAsync(key->d_node_cnt, 0, nodes * sizeof(unsigned), st));
        key->node_ctas = cap;
    }
    TallyPeer P = key->peer;
    if (!collective || P.world <= 1) { P.world = 1; P.rank = 0; }
    else { key->peer.epoch += 1; P.epoch = key->peer.epoch; }
    bool done = false;
    if constexpr (UL<C, 1>::SUPPORTED) if (key->eng >= 2 && key->has_u) {
        k_tally<C, 4><<<(unsigned)ctas, C::THREADS, UL<C, 1>::SMEM_BYTES, st>>>(key->dev, d_c, count, d_out, key->d_nodes, key->d_node_cnt, P);
        done = true;
    }
    if (done) {}
    else if (key->eng >= 1) k_tally<C, 1><<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, st>>>(key->dev, d_c, count, d_out, key->d_nodes, key->d_node_cnt, P);
    else k_tally<C, 0><<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, st>>>(key->dev, d_c, count, d_out, key->d_nodes, key->d_node_cnt, P);
    count_launch();
    return cudaGetLastError();
}


// ---- witness engine: per-key constants and launcher ------------------------------------------------------

template <class C>
static cudaError_t witness_prepare_cfg(Block28Key* key, u64* d_gchain, bool gchain_ready, cudaStream_t st) {
    const BigInt& n = key->n;
    BigInt n2 = BigInt::mul(n, n);
    int sh = C::KN - (int)n2.bits();
    if (sh & 1) sh -= 1;                                   // chain values carry 2^(sh/2)
    BigInt Nt = BigInt::shl(n2, sh);
    BigInt mu = BigInt::div(BigInt::pow2(2 * (size_t)C::BETA), Nt);
    std::vector<int> e_mu, e_nt, e_ntu(C::ENTRY4 * 4, 0), e_one, all;
    to_entry<C>(mu, e_mu); to_entry<C>(Nt, e_nt); to_entry<C>(BigInt::pow2(sh / 2), e_one);
    for (int p = 0; p < C::L; p++) e_ntu[(p / C::BL) * C::CH * 4 + (p % C::BL)] = (int)Nt.bits_at((size_t)W * p, W);
    all.insert(all.end(), e_mu.begin(), e_mu.end());
    all.insert(all.end(), e_nt.begin(), e_nt.end());
    all.insert(all.end(), e_ntu.begin(), e_ntu.end());     // in the slot the value engine uses for 2^sh
    std::vector<int> r_mu, r_nt;
    to_rtab<C>(e_mu, r_mu); to_rtab<C>(e_nt, r_nt);
    all.insert(all.end(), r_mu.begin(), r_mu.end());
    all.insert(all.end(), r_nt.begin(), r_nt.end());
    CUW(cudaMalloc(&key->d_wconsts, all.size() * sizeof(int)));
    CUW(cudaMemcpyAsync(key->d_wconsts, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CUW(cudaMalloc(&key->d_one_s, e_one.size() * sizeof(int)));
    CUW(cudaMemcpyAsync(key->d_one_s, e_one.data(), e_one.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    const int wo = key->dev.words_out, wi = key->dev.words_in;
    std::vector<u64> cpow(2 * (size_t)wo);
    { u64 c = PB200_DIGEST_C; for (size_t j = 0; j < cpow.size(); j++) { cpow[j] = c; c *= PB200_DIGEST_C; } }
    CUW(cudaMalloc(&key->d_cpow, cpow.size() * sizeof(u64)));
    CUW(cudaMemcpyAsync(key->d_cpow, cpow.data(), cpow.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    std::vector<u64> nw(wi);
    n.to_u64_le(nw.data(), wi);
    CUW(cudaMalloc(&key->d_nwords, nw.size() * sizeof(u64)));
    CUW(cudaMemcpyAsync(key->d_nwords, nw.data(), nw.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    CUW(cudaMalloc(&key->d_gtab, (size_t)key->n_bits * C::ENTRY4 * sizeof(int4)));
    // 2^(28(L-2)) / Nt_w from the top 52 bits of Nt_w
    BigInt top = BigInt::shr(Nt, (size_t)W * (C::L - 2) - 20);
    double topd = (double)top.bits_at(32, 32) * 4294967296.0 + (double)top.bits_at(0, 32);
    key->wdev = key->dev; key->wdev.consts = key->d_wconsts; key->wdev.sh = sh;
    WitDev& Wd = key->wit;
    Wd.gtab = key->d_gtab; Wd.one_s = key->d_one_s; Wd.cpow = key->d_cpow; Wd.n_words = key->d_nwords;
    Wd.exp_bits = (int)n.bits(); Wd.sh = sh; Wd.inv = 1048576.0 / topd;
    CUW((cudaFuncSetAttribute(k_witness<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_W_BYTES)));
    CUW((cudaFuncSetAttribute(k_gchain_w<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_W_BYTES)));
    CUW((cudaFuncSetAttribute(k_add_w<C, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_W_BYTES)));
    if constexpr (UL<C, 1, true>::SUPPORTED) {
        // the same step on the tcgen05 phases: constants of the witness modulus + their Toeplitz core-matrix tables
        typedef UL<C, 1, true> U;
        std::vector<signed char> img(U::KEY_BYTES, 0);
        memcpy(img.data(), all.data(), 3 * C::ENTRY4 * 16);
        std::vector<signed char> k7(C::K7);
        to_k7<C>(e_mu, k7);
        umma_cm_table<C, 1>(k7.data(), true, img.data() + (U::OFF_CMH - U::OFF_CONST));
        to_k7<C>(e_nt, k7);
        umma_cm_table<C, 1>(k7.data(), false, img.data() + (U::OFF_CML - U::OFF_CONST));
        CUW(cudaMalloc(&key->d_wuconsts, img.size()));
        CUW(cudaMemcpyAsync(key->d_wuconsts, img.data(), img.size(), cudaMemcpyHostToDevice, st));
        CUW(cudaStreamSynchronize(st));
        key->wdev.uconsts = key->d_wuconsts;
        key->has_wu = !getenv("PB200_NO_UMMA_WITNESS");
        CUW((cudaFuncSetAttribute(k_witness<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U::SMEM_BYTES)));
        CUW((cudaFuncSetAttribute(k_gchain_w<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U::SMEM_BYTES)));
        CUW((cudaFuncSetAttribute(k_add_w<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U::SMEM_BYTES)));
    }
    if (gchain_ready) {        // g-chain records already produced (simple64): only convert them into table entries
        k_wtab<C><<<(key->n_bits + 63) / 64, 64, 0, st>>>(key->d_gwords, wi, d_gchain, wo, (int)key->n_bits, sh / 2, key->d_gtab);
    } else {                   // produce records and table entries with the witness engine itself
        bool done = false;
        if constexpr (UL<C, 1, true>::SUPPORTED) if (key->has_wu && key->eng >= 2) {
            k_gchain_w<C, 1><<<1, C::THREADS, UL<C, 1, true>::SMEM_BYTES, st>>>(key->wdev, key->wit, key->d_gwords, (int)key->n_bits, d_gchain, key->d_gtab);
            done = true;
        }
        if (!done) k_gchain_w<C, 0><<<1, C::THREADS, C::SMEM_W_BYTES, st>>>(key->wdev, key->wit, key->d_gwords, (int)key->n_bits, d_gchain, key->d_gtab);
    }
    count_launch();
    CUW(cudaGetLastError());
    CUW(cudaStreamSynchronize(st));                        // the staging vectors above go out of scope
    key->wit_ready = true;
    return cudaSuccess;
}

template <class C>
static cudaError_t witness_cfg(Block28Key* key, const u64* d_m, const u64* d_r, size_t count, u64* d_c, u64* d_records,
                               const u64* d_offsets, u64* d_digest, cudaStream_t st) {
    size_t ctas = (count + 31) / 32;
    if (ctas > key->wscratch_ctas) {
        if (key->d_wscratch) cudaFree(key->d_wscratch);
        key->d_wscratch = nullptr; key->wscratch_ctas = 0;
        CUW(cudaMalloc(&key->d_wscratch, ctas * 2 * C::VAL4 * sizeof(int4)));
        key->wscratch_ctas = ctas;
    }
    WitIO A;
    A.m = d_m; A.r = d_r; A.count = count; A.c_out = d_c; A.records = d_records; A.offsets = d_offsets; A.digest = d_digest;
    A.scratch = key->d_wscratch;
    bool done = false;
    if constexpr (UL<C, 1, true>::SUPPORTED) if (key->has_wu && key->eng >= 2) {
        k_witness<C, 1><<<(unsigned)ctas, C::THREADS, UL<C, 1, true>::SMEM_BYTES, st>>>(key->wdev, key->wit, A);
        done = true;
    }
    if (!done) k_witness<C, 0><<<(unsigned)ctas, C::THREADS, C::SMEM_W_BYTES, st>>>(key->wdev, key->wit, A);
    count_launch();
    return cudaGetLastError();
}

// Compiled configurations <G warps, BL digits per block>: L = G*BL digits of 28 bits; n^2 must fit KN - 2 bits.
typedef Cfg<4, 19> Cfg1024;    // L =  76: n^2 up to 2102 bits (|n| <= 1024 and the reference's default sizes)
typedef Cfg<8, 19> Cfg2048;    // L = 152: n^2 up to 4230 bits (|n| <= 2048)
typedef Cfg<16, 14> Cfg3072;   // L = 224: n^2 up to 6246 bits (|n| <= 3072)
typedef Cfg<16, 19> Cfg4096;   // L = 304: n^2 up to 8486 bits (|n| <= 4096)

template <class C>
static bool covers(uint32_t n_bits) { return 2 * (size_t)n_bits <= (size_t)C::KN - 2 && n_bits % 8 == 0; }

// layout constants of a block28u variant (host only, no device needed): the CPU model of the tcgen05 data path in
// tests/test_umma_layout.py checks its own derivation against these
template <class C, int LG, bool WIT>
static void ul_fill(int* out) {
    typedef UL<C, LG, WIT> U;
    const int v[20] = {(int)U::SUPPORTED, U::RD, U::TCOLS, U::TN, U::NT_H, U::NT_L, U::FRONT, U::BACK, U::KOFF, U::CHB, U::A_BYTES, U::Z0_H, U::Z0_L,
                       U::NCM_H, U::NCM_L, (int)U::SMEM_BYTES, U::CTAS_PER_SM, U::TMEM_COLS, U::GAP, U::P_BASE_H};
    for (int i = 0; i < 20; i++) out[i] = v[i];
}
template <class C>
static bool ul_cfg(int lg, int wit, int* out) {
    if (lg == 1 && !wit) ul_fill<C, 1, false>(out);
    else if (lg == 1 && wit) ul_fill<C, 1, true>(out);
    else if (lg == 2 && !wit) ul_fill<C, 2, false>(out);
    else return false;
    return true;
}
bool block28_umma_layout(int G, int BL, int lg, int wit, int* out) {
    if (G == 4 && BL == 19) return ul_cfg<Cfg1024>(lg, wit, out);
    if (G == 8 && BL == 19) return ul_cfg<Cfg2048>(lg, wit, out);
    if (G == 16 && BL == 14) return ul_cfg<Cfg3072>(lg, wit, out);
    if (G == 16 && BL == 19) return ul_cfg<Cfg4096>(lg, wit, out);
    return false;
}

Block28Key* block28_create(const BigInt& n, const BigInt& g, uint32_t n_bits, int device, cudaStream_t st,
                           std::string* why, cudaError_t* cuda_err) {
    *cuda_err = cudaSuccess;
    if (covers<Cfg1024>(n_bits)) return create_cfg<Cfg1024>(n, g, n_bits, device, st, cuda_err);
    if (covers<Cfg2048>(n_bits)) return create_cfg<Cfg2048>(n, g, n_bits, device, st, cuda_err);
    if (covers<Cfg3072>(n_bits)) return create_cfg<Cfg3072>(n, g, n_bits, device, st, cuda_err);
    if (covers<Cfg4096>(n_bits)) return create_cfg<Cfg4096>(n, g, n_bits, device, st, cuda_err);
    if (why) *why = "block28: no compiled configuration covers this key size";
    return nullptr;
}
void block28_destroy(Block28Key* key) {
    if (!key) return;
    if (key->d_consts) cudaFree(key->d_consts);
    if (key->d_uconsts) cudaFree(key->d_uconsts);
    if (key->d_uconsts2) cudaFree(key->d_uconsts2);
    if (key->d_ops) cudaFree(key->d_ops);
    if (key->d_tg) cudaFree(key->d_tg);
    if (key->d_nentry) cudaFree(key->d_nentry);
    if (key->d_gwords) cudaFree(key->d_gwords);
    if (key->d_scratch) cudaFree(key->d_scratch);
    if (key->d_slot_masks) cudaFree(key->d_slot_masks);
    if (key->d_pow_ops) cudaFree(key->d_pow_ops);
    if (key->d_nodes) cudaFree(key->d_nodes);
    if (key->d_node_cnt) cudaFree(key->d_node_cnt);
    for (int i = 0; i < TALLY_MAXW; i++) if (key->ipc_opened[i]) cudaIpcCloseMemHandle(key->ipc_opened[i]);
    if (key->d_mail) cudaFree(key->d_mail);
    if (key->d_wconsts) cudaFree(key->d_wconsts);
    if (key->d_wuconsts) cudaFree(key->d_wuconsts);
    if (key->d_gtab) cudaFree(key->d_gtab);
    if (key->d_one_s) cudaFree(key->d_one_s);
    if (key->d_cpow) cudaFree(key->d_cpow);
    if (key->d_nwords) cudaFree(key->d_nwords);
    if (key->d_wscratch) cudaFree(key->d_wscratch);
    delete key;
}
const char* block28_name(const Block28Key* key) { return key->name.c_str(); }
// eng: 0 = every phase on the IMAD pipe, 1 = constant-operand phases on mma.sync, 2 = on tcgen05 with 32 ciphertexts per CTA,
// 3 = on tcgen05 with 64 per CTA (each falls back to the next lower one this configuration has), -1 = the fastest available (2)
int block28_set_engine(Block28Key* key, int eng) {
    if (eng < 0) eng = getenv("PB200_NO_UMMA") ? 1 : (getenv("PB200_UMMA_LG2") ? 3 : 2);      // measured: 134.9 k enc/s (2) against 129.9 k (3)
    if (eng == 3 && !key->has_u2) eng = 2;
    if (eng == 2 && !key->has_u) eng = 1;
    key->eng = eng;
    key->name = std::string(eng == 3 ? "block28u2<" : eng == 2 ? "block28u<" : eng == 1 ? "block28t<" : "block28<") + std::to_string(key->G) + "," +
                std::to_string(key->BL) + ">";
    return eng;
}
bool block28_has_umma(const Block28Key* key) { return key->has_u; }
bool block28_has_umma2(const Block28Key* key) { return key->has_u2; }
void block28_chain_counts(const Block28Key* key, uint64_t* n_sqr, uint64_t* n_mul) { *n_sqr = key->n_sqr; *n_mul = key->n_mul; }
cudaError_t block28_encrypt(Block28Key* key, const u64* d_m, const u64* d_r, size_t count, u64* d_c, cudaStream_t st) {
    if (key->G == 4) return encrypt_cfg<Cfg1024>(key, d_m, d_r, count, d_c, st);
    if (key->G == 8) return encrypt_cfg<Cfg2048>(key, d_m, d_r, count, d_c, st);
    if (key->BL == 14) return encrypt_cfg<Cfg3072>(key, d_m, d_r, count, d_c, st);
    return encrypt_cfg<Cfg4096>(key, d_m, d_r, count, d_c, st);
}
cudaError_t block28_pow_prepare(Block28Key* key, const BigInt& e, cudaStream_t st) {
    std::vector<int2> ops;
    const int first_idx = window_schedule(e, ops);
    if (key->d_pow_ops) { cudaFree(key->d_pow_ops); key->d_pow_ops = nullptr; }
    CUW(cudaMalloc(&key->d_pow_ops, (ops.size() + 1) * sizeof(int2)));
    if (!ops.empty()) CUW(cudaMemcpyAsync(key->d_pow_ops, ops.data(), ops.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    CUW(cudaStreamSynchronize(st));
    key->pow.ops = key->d_pow_ops; key->pow.n_ops = (int)ops.size(); key->pow.first_idx = first_idx;
    key->pow_ready = true;
    return cudaSuccess;
}
cudaError_t block28_pow(Block28Key* key, const u64* d_base, int base_words, size_t count, u64* d_out, cudaStream_t st) {
    if (!count) return cudaSuccess;
    if (!key->pow_ready) return cudaErrorInvalidValue;
    if (key->G == 4) return pow_cfg<Cfg1024>(key, d_base, base_words, count, d_out, st);
    if (key->G == 8) return pow_cfg<Cfg2048>(key, d_base, base_words, count, d_out, st);
    if (key->BL == 14) return pow_cfg<Cfg3072>(key, d_base, base_words, count, d_out, st);
    return pow_cfg<Cfg4096>(key, d_base, base_words, count, d_out, st);
}

static cudaError_t tally_dispatch(Block28Key* key, const u64* d_c, size_t count, u64* d_out, bool collective, cudaStream_t st) {
    if (key->G == 4) return tally_cfg<Cfg1024>(key, d_c, count, d_out, collective, st);
    if (key->G == 8) return tally_cfg<Cfg2048>(key, d_c, count, d_out, collective, st);
    if (key->BL == 14) return tally_cfg<Cfg3072>(key, d_c, count, d_out, collective, st);
    return tally_cfg<Cfg4096>(key, d_c, count, d_out, collective, st);
}
cudaError_t block28_tally(Block28Key* key, const u64* d_c, size_t count, u64* d_out, cudaStream_t st) {
    return tally_dispatch(key, d_c, count, d_out, false, st);
}
// collective over the connected peer group: every rank calls it once per tally, in the same order
cudaError_t block28_tally_peer(Block28Key* key, const u64* d_c, size_t count, u64* d_out, cudaStream_t st) {
    return tally_dispatch(key, d_c, count, d_out, true, st);
}
int block28_peer_world(const Block28Key* key) { return key->peer.world; }

// the peer-visible mailbox of this key (allocated on first use, zeroed)
cudaError_t block28_mailbox(Block28Key* key, u64** d_mail, cudaStream_t st) {
    if (!key->d_mail) {
        CUW(cudaMalloc(&key->d_mail, MAIL_U64 * sizeof(u64)));
        CUW(cudaMemsetAsync(key->d_mail, 0, MAIL_U64 * sizeof(u64), st));
        CUW(cudaStreamSynchronize(st));
    }
    *d_mail = key->d_mail;
    return cudaSuccess;
}
size_t block28_mailbox_bytes() { return MAIL_U64 * sizeof(u64); }
int block28_max_world() { return TALLY_MAXW; }
// mail[i]: rank i's mailbox as addressable from this key's device (mail[rank] must be this key's own); opened[i]: non-null when
// the pointer came from cudaIpcOpenMemHandle and must be closed with the key
cudaError_t block28_tally_peer_connect(Block28Key* key, int rank, int world, u64* const* mail, void* const* opened, int* d_flags) {
    if (world < 1 || world > TALLY_MAXW || rank < 0 || rank >= world) return cudaErrorInvalidValue;
    for (int i = 0; i < TALLY_MAXW; i++) {
        if (key->ipc_opened[i]) { cudaIpcCloseMemHandle(key->ipc_opened[i]); key->ipc_opened[i] = nullptr; }
        key->peer.mail[i] = i < world ? mail[i] : nullptr;
        if (opened && i < world) key->ipc_opened[i] = opened[i];
    }
    key->peer.world = world; key->peer.rank = rank; key->peer.epoch = 0; key->peer.flags = d_flags;
    // a reconnect restarts the epochs: the local mailbox flags must restart with them
    if (key->d_mail) CUW(cudaMemset(key->d_mail + MAIL_FLAG_OFF, 0, 2 * TALLY_MAXW * sizeof(u64)));
    return cudaSuccess;
}


template <class C, int ENG>
static cudaError_t debug_launch(Block28Key* key, const int4* d_v, const int4* d_y, int reps, int4* d_vout, int4* d_t, unsigned* d_rows, cudaStream_t st) {
    CUW((cudaFuncSetAttribute(k_mulmod_dbg<C, ENG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)View<C, ENG>::BYTES)));
    k_mulmod_dbg<C, ENG><<<1, View<C, ENG>::THREADS, View<C, ENG>::BYTES, st>>>(key->dev, d_v, d_y, reps, d_vout, d_t, d_rows);
    count_launch();
    return cudaGetLastError();
}
template <class C>
static cudaError_t debug_cfg(Block28Key* key, int eng, const int* h_v, const int* h_y, int reps, int* h_vout, int* h_t, unsigned* h_rows, cudaStream_t st) {
    const size_t vb = (size_t)C::VAL4 * 16;
    int4 *d_v = nullptr, *d_y = nullptr, *d_vout = nullptr, *d_t = nullptr; unsigned* d_rows = nullptr;
    cudaError_t e = cudaSuccess;
    auto run = [&]() -> cudaError_t {
        CUW(cudaMalloc(&d_v, vb)); CUW(cudaMalloc(&d_y, vb)); CUW(cudaMalloc(&d_vout, vb)); CUW(cudaMalloc(&d_t, 2 * vb));
        CUW(cudaMalloc(&d_rows, (size_t)32 * C::L * 4));
        CUW(cudaMemset(d_rows, 0, (size_t)32 * C::L * 4));
        CUW(cudaMemcpy(d_v, h_v, vb, cudaMemcpyHostToDevice));
        if (h_y) CUW(cudaMemcpy(d_y, h_y, vb, cudaMemcpyHostToDevice));
        cudaError_t r = cudaErrorInvalidValue;
        if (eng == 3) { if constexpr (UL<C, 1>::SUPPORTED && UL<C, 2>::SUPPORTED) { if (key->has_u2) r = debug_launch<C, 3>(key, d_v, h_y ? d_y : nullptr, reps, d_vout, h_t ? d_t : nullptr, d_rows, st); } }
        else if (eng == 2) { if constexpr (UL<C, 1>::SUPPORTED) { if (key->has_u) r = debug_launch<C, 2>(key, d_v, h_y ? d_y : nullptr, reps, d_vout, h_t ? d_t : nullptr, d_rows, st); } }
        else if (eng == 1) r = debug_launch<C, 1>(key, d_v, h_y ? d_y : nullptr, reps, d_vout, h_t ? d_t : nullptr, d_rows, st);
        else if (eng == 0) r = debug_launch<C, 0>(key, d_v, h_y ? d_y : nullptr, reps, d_vout, h_t ? d_t : nullptr, d_rows, st);
        CUW(r);
        CUW(cudaStreamSynchronize(st));
        if (h_t) CUW(cudaMemcpy(h_t, d_t, 2 * vb, cudaMemcpyDeviceToHost));
        else {
            CUW(cudaMemcpy(h_vout, d_vout, vb, cudaMemcpyDeviceToHost));
            if (h_rows) CUW(cudaMemcpy(h_rows, d_rows, (size_t)32 * C::L * 4, cudaMemcpyDeviceToHost));
        }
        return cudaSuccess;
    };
    e = run();
    cudaFree(d_v); cudaFree(d_y); cudaFree(d_vout); cudaFree(d_t); cudaFree(d_rows);
    return e;
}
// eng: 0 block28, 1 block28t, 2 block28u.  h_t non-null: phase A only, the 2L-digit product; else `reps` multiplications, V and q-hat rows
cudaError_t block28_debug_mulmod(Block28Key* key, int eng, const int* h_v, const int* h_y, int reps, int* h_vout, int* h_t, unsigned* h_rows,
                                 cudaStream_t st) {
    if (key->G == 4) return debug_cfg<Cfg1024>(key, eng, h_v, h_y, reps, h_vout, h_t, h_rows, st);
    if (key->G == 8) return debug_cfg<Cfg2048>(key, eng, h_v, h_y, reps, h_vout, h_t, h_rows, st);
    if (key->BL == 14) return debug_cfg<Cfg3072>(key, eng, h_v, h_y, reps, h_vout, h_t, h_rows, st);
    return debug_cfg<Cfg4096>(key, eng, h_v, h_y, reps, h_vout, h_t, h_rows, st);
}
void block28_shape(const Block28Key* key, int* G, int* BL) { *G = key->G; *BL = key->BL; }


template <class C, int ENG>
static cudaError_t time_launch(Block28Key* key, const int4* d_v, int ctas, int reps, int stagger, long long* d_cyc, unsigned* d_cnt, cudaStream_t st) {
    CUW((cudaFuncSetAttribute(k_mulmod_time<C, ENG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)View<C, ENG>::BYTES)));
    k_mulmod_time<C, ENG><<<ctas, View<C, ENG>::THREADS, View<C, ENG>::BYTES, st>>>(key->dev, d_v, reps, stagger, d_cyc, d_cnt);
    count_launch();
    return cudaGetLastError();
}
// eng 1 / 2; h_cyc: 3 * ctas values
cudaError_t block28_debug_time(Block28Key* key, int eng, const int* h_v, int ctas, int reps, int stagger, long long* h_cyc, cudaStream_t st) {
    if (key->G != 8) return cudaErrorInvalidValue;
    typedef Cfg2048 C;
    const size_t vb = (size_t)C::VAL4 * 16;
    int4* d_v = nullptr; long long* d_cyc = nullptr; unsigned* d_cnt = nullptr;
    auto run = [&]() -> cudaError_t {
        CUW(cudaMalloc(&d_v, vb)); CUW(cudaMalloc(&d_cyc, (size_t)ctas * 3 * sizeof(long long)));
        CUW(cudaMalloc(&d_cnt, 1024 * sizeof(unsigned))); CUW(cudaMemset(d_cnt, 0, 1024 * sizeof(unsigned)));
        CUW(cudaMemcpy(d_v, h_v, vb, cudaMemcpyHostToDevice));
        cudaError_t r = cudaErrorInvalidValue;
        if (eng == 3) { if (key->has_u2) r = time_launch<C, 3>(key, d_v, ctas, reps, stagger, d_cyc, d_cnt, st); }
        else if (eng == 2) { if (key->has_u) r = time_launch<C, 2>(key, d_v, ctas, reps, stagger, d_cyc, d_cnt, st); }
        else if (eng == 1) r = time_launch<C, 1>(key, d_v, ctas, reps, stagger, d_cyc, d_cnt, st);
        CUW(r);
        CUW(cudaStreamSynchronize(st));
        CUW(cudaMemcpy(h_cyc, d_cyc, (size_t)ctas * 3 * sizeof(long long), cudaMemcpyDeviceToHost));
        return cudaSuccess;
    };
    cudaError_t e = run();
    cudaFree(d_v); cudaFree(d_cyc); cudaFree(d_cnt);
    return e;
}

template <class C>
static cudaError_t add_cfg(Block28Key* key, const u64* d_c1, const u64* d_c2, int c_words, size_t count, u64* d_out, u64* d_q,
                           int* d_flags, cudaStream_t st) {
    size_t ctas = (count + 31) / 32, cap = (size_t)C::CTAS_PER_SM * key->sms;
    if (ctas > cap) ctas = cap;
    bool done = false;
    if constexpr (UL<C, 1, true>::SUPPORTED) if (key->has_wu && key->eng >= 2) {
        k_add_w<C, 1><<<(unsigned)ctas, C::THREADS, UL<C, 1, true>::SMEM_BYTES, st>>>(key->wdev, key->wit, d_c1, d_c2, c_words, count, d_out, d_q, d_flags);
        done = true;
    }
    if (!done) k_add_w<C, 0><<<(unsigned)ctas, C::THREADS, C::SMEM_W_BYTES, st>>>(key->wdev, key->wit, d_c1, d_c2, c_words, count, d_out, d_q, d_flags);
    count_launch();
    return cudaGetLastError();
}

// The witness engine needs canonical chain values below n^2 from the first step on: n must fill its declared width
// (then g, r < 2^n_bits <= 2n <= n^2).  Other keys keep the simple64 witness path.
bool block28_witness_supported(const Block28Key* key) { return key->eng >= 1 && key->n.bits() == key->n_bits && key->n_bits >= 8; }
cudaError_t block28_witness_prepare(Block28Key* key, u64* d_gchain, bool gchain_ready, cudaStream_t st) {
    if (key->wit_ready) return cudaSuccess;
    try {
        if (key->G == 4) return witness_prepare_cfg<Cfg1024>(key, d_gchain, gchain_ready, st);
        if (key->G == 8) return witness_prepare_cfg<Cfg2048>(key, d_gchain, gchain_ready, st);
        if (key->BL == 14) return witness_prepare_cfg<Cfg3072>(key, d_gchain, gchain_ready, st);
        return witness_prepare_cfg<Cfg4096>(key, d_gchain, gchain_ready, st);
    } catch (const std::exception&) { return cudaErrorInvalidValue; }
}
cudaError_t block28_witness(Block28Key* key, const u64* d_m, const u64* d_r, size_t count, u64* d_c, u64* d_records,
                            const u64* d_offsets, u64* d_digest, cudaStream_t st) {
    if (!count) return cudaSuccess;
    if (key->G == 4) return witness_cfg<Cfg1024>(key, d_m, d_r, count, d_c, d_records, d_offsets, d_digest, st);
    if (key->G == 8) return witness_cfg<Cfg2048>(key, d_m, d_r, count, d_c, d_records, d_offsets, d_digest, st);
    if (key->BL == 14) return witness_cfg<Cfg3072>(key, d_m, d_r, count, d_c, d_records, d_offsets, d_digest, st);
    return witness_cfg<Cfg4096>(key, d_m, d_r, count, d_c, d_records, d_offsets, d_digest, st);
}

cudaError_t block28_add(Block28Key* key, const u64* d_c1, const u64* d_c2, int c_words, size_t count, u64* d_out, u64* d_q,
                        int* d_flags, cudaStream_t st) {
    if (!count) return cudaSuccess;
    if (key->G == 4) return add_cfg<Cfg1024>(key, d_c1, d_c2, c_words, count, d_out, d_q, d_flags, st);
    if (key->G == 8) return add_cfg<Cfg2048>(key, d_c1, d_c2, c_words, count, d_out, d_q, d_flags, st);
    if (key->BL == 14) return add_cfg<Cfg3072>(key, d_c1, d_c2, c_words, count, d_out, d_q, d_flags, st);
    return add_cfg<Cfg4096>(key, d_c1, d_c2, c_words, count, d_out, d_q, d_flags, st);
}

}  // namespace pb200
