// cells.cu — K4: the advice-cell values BigUintChip assigns around one mul_mod, produced on the GPU.
//
// Replaces the witness-cell arithmetic of biguint-halo2 that PaillierChip::encrypt/add drive
// (/root/reference/src/paillier.rs:39-45,51,55,57,81; semantics restated in SURVEY.md Appendix A):
//   k_cells_mulmod : one mul_mod(a, b, n^2) group (A.4 + A.6): q limbs, rem limbs (each followed by its range-check
//                    chunks, A.1), the no-carry columns ab and q*n^2 (A.2), the sums q*n^2 + rem, and per column of
//                    is_equal_muled the carry / cs / q_acc / mod_acc values with the carry's range-check chunks,
//                    then the eq flag — in the order the chip assigns them;
//   k_cells_assign : assign_integer (A.1) of arbitrary values: limb, chunks [, shifted top chunk];
//   k_cells_n2     : square(n) + refresh (A.2, A.3) — per key.
// Every cell is a BN254 Fr element written as 4 little-endian u64 words: canonical integer (the values never wrap
// the field: < 2^183) or Montgomery form x * 2^256 mod p (halo2curves' in-memory representation).
//
// This is HBM-bound byte work: a mul_mod group at |n| = 2048 / 64-bit limbs / 15-bit lookups is 2 546 cells = 81 KB
// of output for 8 K limb products, so the kernel is organised around coalesced 32-byte cell stores (thread c writes
// cell c), with the group's primary values (limbs, columns, carries) staged in shared memory.
#include "engine.hpp"
#include "cells.hpp"

namespace pb200 {

typedef unsigned __int128 u128c;

struct W4 { u64 w[4]; };

__device__ __forceinline__ W4 w4_zero() { W4 r; r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0; return r; }
__device__ __forceinline__ W4 w4_from128(u128c v) { W4 r; r.w[0] = (u64)v; r.w[1] = (u64)(v >> 64); r.w[2] = r.w[3] = 0; return r; }
__device__ __forceinline__ void w4_add(W4& a, const W4& b) {
    u128c c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { c += (u128c)a.w[i] + b.w[i]; a.w[i] = (u64)c; c >>= 64; }
}
__device__ __forceinline__ void w4_sub(W4& a, const W4& b) {      // two's complement
    u64 borrow = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 ai = a.w[i], bi = b.w[i];
        u64 d = ai - bi, b1 = ai < bi;
        u64 d2 = d - borrow, b2 = d < borrow;
        a.w[i] = d2; borrow = b1 | b2;
    }
}
// acc += a * b, a and b up to 128 bits (hi words zero for limbs of <= 64 bits)
__device__ __forceinline__ void w4_mac(W4& acc, u64 a0, u64 a1, u64 b0, u64 b1, bool wide) {
    u128c p = (u128c)a0 * b0;
    u128c c = (u128c)acc.w[0] + (u64)p; acc.w[0] = (u64)c; c >>= 64;
    c += (u128c)acc.w[1] + (u64)(p >> 64); acc.w[1] = (u64)c; c >>= 64;
    c += acc.w[2]; acc.w[2] = (u64)c; c >>= 64;
    acc.w[3] += (u64)c;
    if (wide) {
        u128c p1 = (u128c)a0 * b1, p2 = (u128c)a1 * b0, p3 = (u128c)a1 * b1;
        c = (u128c)acc.w[1] + (u64)p1 + (u64)p2; acc.w[1] = (u64)c; c >>= 64;
        c += (u128c)acc.w[2] + (u64)(p1 >> 64) + (u64)(p2 >> 64) + (u64)p3; acc.w[2] = (u64)c; c >>= 64;
        acc.w[3] += (u64)c + (u64)(p3 >> 64);
    }
}
__device__ __forceinline__ W4 w4_shr(const W4& a, int s) {        // logical, 0 <= s < 256
    W4 r; const int ws = s >> 6, bs = s & 63;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u64 lo = i + ws < 4 ? a.w[i + ws] : 0, hi = i + ws + 1 < 4 ? a.w[i + ws + 1] : 0;
        r.w[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
    return r;
}
__device__ __forceinline__ W4 w4_shl(const W4& a, int s) {        // 0 <= s < 64
    W4 r;
#pragma unroll
    for (int i = 3; i >= 0; i--) r.w[i] = s ? (a.w[i] << s) | (i ? a.w[i - 1] >> (64 - s) : 0) : a.w[i];
    return r;
}
__device__ __forceinline__ W4 w4_low(const W4& a, int bits) {     // a mod 2^bits, bits < 256
    W4 r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int lo = 64 * i;
        r.w[i] = bits >= lo + 64 ? a.w[i] : (bits <= lo ? 0 : a.w[i] & ((1ull << (bits - lo)) - 1));
    }
    return r;
}
__device__ __forceinline__ bool w4_eq(const W4& a, const W4& b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3]; }

// bits [lo, lo+cnt) of a little-endian word array, cnt <= 128
__device__ __forceinline__ u128c bits_of(const u64* v, int nwords, int lo, int cnt) {
    const int wi = lo >> 6, sh = lo & 63;
    u64 w0 = wi < nwords ? v[wi] : 0, w1 = wi + 1 < nwords ? v[wi + 1] : 0, w2 = wi + 2 < nwords ? v[wi + 2] : 0;
    u64 a = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
    u64 b = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
    u128c r = ((u128c)b << 64) | a;
    if (cnt < 128) r &= (((u128c)1) << cnt) - 1;
    return r;
}

// ---- BN254 Fr, Montgomery form (R = 2^256) ----------------------------------------------------------------
__device__ __constant__ u64 FR_P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
__device__ __constant__ u64 FR_R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
#define FR_INV 0xc2e1f593efffffffull

// x * R mod p for x < p (CIOS Montgomery product of x and R^2)
__device__ __forceinline__ W4 fr_to_mont(const W4& x) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u128c c = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { c += (u128c)x.w[j] * FR_R2[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
        const u64 m = t[0] * FR_INV;
        c = (u128c)m * FR_P[0] + t[0]; c >>= 64;
#pragma unroll
        for (int j = 1; j < 4; j++) { c += (u128c)m * FR_P[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    W4 r; r.w[0] = t[0]; r.w[1] = t[1]; r.w[2] = t[2]; r.w[3] = t[3];
    // conditional subtraction
    bool ge = t[4] != 0;
    if (!ge) {
        ge = true;
#pragma unroll
        for (int i = 3; i >= 0; i--) if (r.w[i] != FR_P[i]) { ge = r.w[i] > FR_P[i]; break; }
    }
    if (ge) { W4 p; p.w[0] = FR_P[0]; p.w[1] = FR_P[1]; p.w[2] = FR_P[2]; p.w[3] = FR_P[3]; w4_sub(r, p); }
    return r;
}

// ---- Montgomery conversion on 32-bit words --------------------------------------------------------------------------------
// x * R mod p = REDC(x * R^2), CIOS in four 64-bit steps.  The running value T lives in two arrays of 32-bit words whose
// 64-bit pairs sit at even (E) and odd (O) word offsets, T = sum (E_k + O_k) 2^(32k): every 32 x 32 product then lands on an
// aligned register pair and a row of four products is one carry chain, which ptxas compiles to IMAD.WIDE.U32(.X) — about 4.5
// instructions per 64 x 64 product instead of the 17 the compiler makes of unsigned __int128 code.  NX = number of 64-bit words
// of x that can be non-zero (3 for a no-carry column, 2 for a carry).  Modelled word for word in Python before it was written.
__device__ __constant__ unsigned FR_P32[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
__device__ __constant__ unsigned FR_R2_32[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
// t[0..7] (four 64-bit pairs) += a * {b0, b1, b2, b3}, one product per pair, carry chained and carried on through t[8..10]
#define PB200_CHAIN4(t, k, a, b0, b1, b2, b3)                                                                                   \
    asm("mad.lo.cc.u32 %0, %11, %12, %0; madc.hi.cc.u32 %1, %11, %12, %1; madc.lo.cc.u32 %2, %11, %13, %2; madc.hi.cc.u32 %3, %11, %13, %3;" \
        "madc.lo.cc.u32 %4, %11, %14, %4; madc.hi.cc.u32 %5, %11, %14, %5; madc.lo.cc.u32 %6, %11, %15, %6; madc.hi.cc.u32 %7, %11, %15, %7;" \
        "addc.cc.u32 %8, %8, 0; addc.cc.u32 %9, %9, 0; addc.u32 %10, %10, 0;"                                                    \
        : "+r"(t[k]), "+r"(t[k + 1]), "+r"(t[k + 2]), "+r"(t[k + 3]), "+r"(t[k + 4]), "+r"(t[k + 5]), "+r"(t[k + 6]), "+r"(t[k + 7]),   \
          "+r"(t[k + 8]), "+r"(t[k + 9]), "+r"(t[k + 10])                                                                        \
        : "r"(a), "r"(b0), "r"(b1), "r"(b2), "r"(b3))
// T += a (64 bits) * b (8 words) at word offset BASE (even)
#define PB200_ROW(E, O, BASE, a, b)                                                                                              \
    {                                                                                                                            \
        const unsigned a0_ = (unsigned)(a), a1_ = (unsigned)((a) >> 32);                                                         \
        PB200_CHAIN4(E, BASE, a0_, b[0], b[2], b[4], b[6]);                                                                      \
        PB200_CHAIN4(O, BASE + 1, a0_, b[1], b[3], b[5], b[7]);                                                                  \
        PB200_CHAIN4(O, BASE + 1, a1_, b[0], b[2], b[4], b[6]);                                                                  \
        PB200_CHAIN4(E, BASE + 2, a1_, b[1], b[3], b[5], b[7]);                                                                  \
    }
template <int NX>
__device__ __forceinline__ W4 fr_to_mont32(const W4& x) {
    unsigned E[20], O[20];
#pragma unroll
    for (int k = 0; k < 20; k++) { E[k] = 0; O[k] = 0; }
    unsigned p[8], r2[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { p[k] = FR_P32[k]; r2[k] = FR_R2_32[k]; }
    u64 cin = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < NX) PB200_ROW(E, O, 2 * i, x.w[i], r2)
        u64 s0 = (u64)E[2 * i] + O[2 * i] + cin;
        u64 s1 = (u64)E[2 * i + 1] + O[2 * i + 1] + (s0 >> 32);
        const u64 ti = (s0 & 0xffffffffull) | (s1 << 32);
        const u64 m = ti * FR_INV;
        PB200_ROW(E, O, 2 * i, m, p)
        s0 = (u64)E[2 * i] + O[2 * i] + cin;
        s1 = (u64)E[2 * i + 1] + O[2 * i + 1] + (s0 >> 32);
        cin = s1 >> 32;
    }
    W4 a, b;
#pragma unroll
    for (int k = 0; k < 4; k++) { a.w[k] = ((u64)E[9 + 2 * k] << 32) | E[8 + 2 * k]; b.w[k] = ((u64)O[9 + 2 * k] << 32) | O[8 + 2 * k]; }
    w4_add(a, b);
    b = w4_zero(); b.w[0] = cin;
    w4_add(a, b);
    // conditional subtraction of p, branch-free
    W4 d = a;
    u64 borrow = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u64 xk = d.w[k], yk = FR_P[k], t = xk - yk, b1 = xk < yk, t2 = t - borrow, b2 = t < borrow;
        d.w[k] = t2; borrow = b1 | b2;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) a.w[k] = borrow ? a.w[k] : d.w[k];
    return a;
}
// x * R mod p for x < 2^(64 NX), NX = 1, 2, 3, branch-free: REDC with the SHORT radix 2^(64 NX) of T = x * K_NX, K_NX = 2^(256 + 64 NX) mod p.
// T < 2^(64 NX) p, so NX reduction rows leave T / 2^(64 NX) < 2p: 2 NX rows of sixteen 32 x 32 products in all (one-word cells: 32
// products instead of the 40 + estimate + correction loops of fr_to_mont_u64; carries: 64 instead of 112; columns: 96 instead of 112)
// and one conditional subtraction.  Not inlined: the conversion pass of the Montgomery kernel is three call sites, not 170 KB of SASS.
__device__ __constant__ unsigned FR_K32[3][8] = {
    {0x7c5fb586u, 0xb4c6edf9u, 0xbfeb93beu, 0x708c8d50u, 0x04f7e0efu, 0x9ffd1de4u, 0x9a392866u, 0x215b02acu},
    {0xef8cfeb9u, 0xb075da81u, 0xa5b6cd8cu, 0xa7f12accu, 0x7957bf7bu, 0x32c47504u, 0x48ffa25eu, 0x03d581d7u},
    {0xc177f51au, 0x5665c3b5u, 0xde75c713u, 0x00e7f02au, 0x2f747168u, 0xb09192e5u, 0xcccdc65du, 0x0621c0bbu}};
template <int NX>
__device__ __noinline__ W4 fr_redc(u64 x0, u64 x1, u64 x2) {
#ifdef PB200_CELLS_FAKE_REDC      // experiment only (tools/build_variant.py): how much of the kernel is the conversion arithmetic
    { W4 f; f.w[0] = x0; f.w[1] = x1; f.w[2] = x2; f.w[3] = NX; return f; }
#endif
    unsigned E[2 * NX + 12], O[2 * NX + 12];
#pragma unroll
    for (int k = 0; k < 2 * NX + 12; k++) { E[k] = 0; O[k] = 0; }
    unsigned p[8], kk[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { p[k] = FR_P32[k]; kk[k] = FR_K32[NX - 1][k]; }
    const u64 xs[3] = {x0, x1, x2};
#pragma unroll
    for (int i = 0; i < NX; i++) PB200_ROW(E, O, 2 * i, xs[i], kk)
    u64 cin = 0;
#pragma unroll
    for (int i = 0; i < NX; i++) {
        u64 s0 = (u64)E[2 * i] + O[2 * i] + cin;
        u64 s1 = (u64)E[2 * i + 1] + O[2 * i + 1] + (s0 >> 32);
        const u64 ti = (s0 & 0xffffffffull) | (s1 << 32);
        const u64 m = ti * FR_INV;
        PB200_ROW(E, O, 2 * i, m, p)
        s0 = (u64)E[2 * i] + O[2 * i] + cin;
        s1 = (u64)E[2 * i + 1] + O[2 * i + 1] + (s0 >> 32);
        cin = s1 >> 32;
    }
    W4 a, b;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        a.w[k] = ((u64)E[2 * NX + 1 + 2 * k] << 32) | E[2 * NX + 2 * k];
        b.w[k] = ((u64)O[2 * NX + 1 + 2 * k] << 32) | O[2 * NX + 2 * k];
    }
    w4_add(a, b);
    b = w4_zero(); b.w[0] = cin;
    w4_add(a, b);
    W4 d = a;
    u64 borrow = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u64 xk = d.w[k], yk = FR_P[k], t = xk - yk, b1 = xk < yk, t2 = t - borrow, b2 = t < borrow;
        d.w[k] = t2; borrow = b1 | b2;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) a.w[k] = borrow ? a.w[k] : d.w[k];
    return a;
}
#undef PB200_ROW
#undef PB200_CHAIN4

// x * R mod p for a one-word x: x * (R mod p) is 5 words, the quotient by p fits one word and is estimated from the top
// 128 bits of the product with a 64-bit reciprocal of p's top word (never low, at most one too high: checked over 2*10^5
// values incl. the extremes), then one conditional correction.  10 word products instead of the 36 of the general path.
__device__ __constant__ u64 FR_R1[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
#define FR_RHO 0xa948e8c4c4740950ull      // floor(2^127 / (p >> 190))
__device__ __forceinline__ W4 fr_to_mont_u64(u64 x) {
    u64 t[5];
    u128c c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { c += (u128c)x * FR_R1[j]; t[j] = (u64)c; c >>= 64; }
    t[4] = (u64)c;
    const u64 X0 = (t[2] >> 62) | (t[3] << 2), X1 = (t[3] >> 62) | (t[4] << 2);
    const u128c e = (u128c)X1 * FR_RHO + __umul64hi(X0, FR_RHO);
    const u64 qh = (u64)(e >> 63);
    u64 m[5];
    c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { c += (u128c)qh * FR_P[j]; m[j] = (u64)c; c >>= 64; }
    m[4] = (u64)c;
    u64 borrow = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const u64 d = t[j] - m[j], b1 = t[j] < m[j], d2 = d - borrow, b2 = d < borrow;
        t[j] = d2; borrow = b1 | b2;
    }
    W4 r; r.w[0] = t[0]; r.w[1] = t[1]; r.w[2] = t[2]; r.w[3] = t[3];
    W4 pp; pp.w[0] = FR_P[0]; pp.w[1] = FR_P[1]; pp.w[2] = FR_P[2]; pp.w[3] = FR_P[3];
    int guard = 0;
    while ((t[4] >> 63) && guard++ < 4) {          // negative: add p
        u128c cc = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { cc += (u128c)r.w[j] + pp.w[j]; r.w[j] = (u64)cc; cc >>= 64; }
        t[4] += (u64)cc;
    }
    for (guard = 0; guard < 4; guard++) {           // >= p: subtract (not expected)
        bool ge = t[4] != 0;
        if (!ge) {
            ge = true;
#pragma unroll
            for (int i = 3; i >= 0; i--) if (r.w[i] != pp.w[i]) { ge = r.w[i] > pp.w[i]; break; }
        }
        if (!ge) break;
        u64 bw = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { const u64 d = r.w[j] - pp.w[j], b1 = r.w[j] < pp.w[j], d2 = d - bw, b2 = d < bw; r.w[j] = d2; bw = b1 | b2; }
        t[4] -= bw;
    }
    return r;
}

// (a + b) mod p for a, b < p, branch-free
__device__ __forceinline__ W4 fr_add(W4 a, const W4& b) {
    w4_add(a, b);                                   // < 2p < 2^255: no carry out of 256 bits
    W4 d = a, pp;
    pp.w[0] = FR_P[0]; pp.w[1] = FR_P[1]; pp.w[2] = FR_P[2]; pp.w[3] = FR_P[3];
    u64 borrow = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u64 x = d.w[i], y = pp.w[i], t = x - y, b1 = x < y, t2 = t - borrow, b2 = t < borrow;
        d.w[i] = t2; borrow = b1 | b2;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) a.w[i] = borrow ? a.w[i] : d.w[i];
    return a;
}

enum CellKind { CK_RAW = 0, CK_SMALL = 1, CK_WIDE = 2 };   // already in output form / one-word value / up to four words
template <bool MONT>
__device__ __forceinline__ void store_cell_k(u64* out, size_t cell, W4 v, int kind) {
    if (MONT) {
        if (kind == CK_SMALL) v = fr_to_mont_u64(v.w[0]);
        else if (kind == CK_WIDE) v = fr_to_mont32<3>(v);          // the wide cells of this kernel are at most 136 bits
    }
    ulonglong4 o; o.x = v.w[0]; o.y = v.w[1]; o.z = v.w[2]; o.w = v.w[3];
    reinterpret_cast<ulonglong4*>(out)[cell] = o;
}

// Montgomery forms of all lookup-chunk values 0 .. 2^lookup_bits - 1 (most cells are range-check chunks): one 32-byte
// gather from an L2-resident table replaces the conversion arithmetic
__global__ void k_mont_table(int n, u64* __restrict__ tab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const W4 v = fr_to_mont_u64((u64)x);
    tab[4 * x] = v.w[0]; tab[4 * x + 1] = v.w[1]; tab[4 * x + 2] = v.w[2]; tab[4 * x + 3] = v.w[3];
}
template <bool MONT>
__device__ __forceinline__ void store_chunk(u64* out, size_t cell, W4 v, const u64* __restrict__ mtab) {
    if (MONT && mtab) {
        const ulonglong2* t = reinterpret_cast<const ulonglong2*>(mtab) + 2 * v.w[0];
        const ulonglong2 a = __ldg(t), b = __ldg(t + 1);
        v.w[0] = a.x; v.w[1] = a.y; v.w[2] = b.x; v.w[3] = b.y;
        store_cell_k<MONT>(out, cell, v, CK_RAW);
    } else if (MONT) store_cell_k<true>(out, cell, fr_redc<1>(v.w[0], 0, 0), CK_RAW);
    else store_cell_k<false>(out, cell, v, CK_RAW);
}

// x * R mod p for a 32-bit x, branch-free and without memory: v * C with C = R mod p (nine words), the quotient by p estimated as
// q-hat = (v * floor(2^32 C / p)) >> 32 — exact or one low, because C / p - M / 2^32 < 2^-32 and v < 2^32 — then v C - q-hat p in [0, 2p)
// and one conditional subtraction.  Eight + eight 32 x 32 products on independent register pairs (even / odd word offsets), no table.
// The range-check chunks are 60 % of the cells: as table gathers they cost two 32-sector L1 wavefront bursts per 32 cells and the
// kernel sat on the LSU data pipe (64 % busy, profiles/ncu_k_cells_r02_mont_first_summary.txt); as arithmetic they cost ~60 instructions.
#define FR_CM32 0x4a474626u          // floor(2^32 (R mod p) / p)
__device__ __forceinline__ void mul8_u32(u64 (&t)[4], unsigned v, const unsigned (&c)[8]) {     // t = v * c mod 2^256 (the 9th word is not needed)
    const u64 e0 = (u64)v * c[0], e1 = (u64)v * c[2], e2 = (u64)v * c[4], e3 = (u64)v * c[6];
    const u64 o0 = (u64)v * c[1], o1 = (u64)v * c[3], o2 = (u64)v * c[5], o3 = (u64)v * c[7];
    // T = E + (O << 32)
    const u64 s0 = o0 << 32, s1 = (o0 >> 32) | (o1 << 32), s2 = (o1 >> 32) | (o2 << 32), s3 = (o2 >> 32) | (o3 << 32);
    asm("add.cc.u64 %0, %4, %8; addc.cc.u64 %1, %5, %9; addc.cc.u64 %2, %6, %10; addc.u64 %3, %7, %11;"
        : "=l"(t[0]), "=l"(t[1]), "=l"(t[2]), "=l"(t[3]) : "l"(e0), "l"(e1), "l"(e2), "l"(e3), "l"(s0), "l"(s1), "l"(s2), "l"(s3));
}
__device__ __forceinline__ W4 fr_mont_u32(unsigned v) {
#ifdef PB200_CELLS_FAKE_CHUNK     // experiment only
    { W4 f = w4_zero(); f.w[0] = v; return f; }
#endif
    unsigned c[8], pw[8];
#pragma unroll
    for (int k = 0; k < 4; k++) { c[2 * k] = (unsigned)FR_R1[k]; c[2 * k + 1] = (unsigned)(FR_R1[k] >> 32); pw[k] = FR_P32[k]; pw[k + 4] = FR_P32[k + 4]; }
    const unsigned qh = __umulhi(v, FR_CM32);
    u64 t[4], u[4];
    mul8_u32(t, v, c);
    mul8_u32(u, qh, pw);
    // r = t - u (mod 2^256; the true value is in [0, 2p) < 2^255), then r - p if that is not negative
    u64 r0, r1, r2, r3, d0, d1, d2, d3, bw;
    asm("sub.cc.u64 %0, %4, %8; subc.cc.u64 %1, %5, %9; subc.cc.u64 %2, %6, %10; subc.u64 %3, %7, %11;"
        : "=l"(r0), "=l"(r1), "=l"(r2), "=l"(r3) : "l"(t[0]), "l"(t[1]), "l"(t[2]), "l"(t[3]), "l"(u[0]), "l"(u[1]), "l"(u[2]), "l"(u[3]));
    asm("sub.cc.u64 %0, %5, %9; subc.cc.u64 %1, %6, %10; subc.cc.u64 %2, %7, %11; subc.cc.u64 %3, %8, %12; subc.u64 %4, 0, 0;"
        : "=l"(d0), "=l"(d1), "=l"(d2), "=l"(d3), "=l"(bw) : "l"(r0), "l"(r1), "l"(r2), "l"(r3), "l"(FR_P[0]), "l"(FR_P[1]), "l"(FR_P[2]), "l"(FR_P[3]));
    W4 o;
    o.w[0] = bw ? r0 : d0; o.w[1] = bw ? r1 : d1; o.w[2] = bw ? r2 : d2; o.w[3] = bw ? r3 : d3;
    return o;
}

// Four chunk cells per lane and pass: the Montgomery forms are 32-byte gathers from the L2-resident table, and with one gather per
// pass the kernel sat on their latency (profiles/ncu_k_cells_r02_mont_summary.txt: long-scoreboard 2.4 stall cycles per issue at 16
// warps per SM).  All eight loads of a pass are issued before the first store.
template <bool MONT>
__device__ __forceinline__ void store_chunks4(u64* out, const size_t (&cell)[4], const u64 (&ch)[4], const bool (&ok)[4], const u64* __restrict__ mtab) {
#ifdef PB200_CELLS_CHUNK_COMPUTE
    if (MONT) {          // A/B variant: computed instead of gathered (measured 2.68 ms against 2.53 ms per 2^16 groups: not the default)
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) store_cell_k<true>(out, cell[u], fr_mont_u32((unsigned)ch[u]), CK_RAW);
        return;
    }
#endif
    if (MONT && mtab) {
        ulonglong2 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const ulonglong2* t = reinterpret_cast<const ulonglong2*>(mtab) + 2 * (ok[u] ? ch[u] : 0);
            a[u] = __ldg(t); b[u] = __ldg(t + 1);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) { ulonglong4 o; o.x = a[u].x; o.y = a[u].y; o.z = b[u].x; o.w = b[u].y; reinterpret_cast<ulonglong4*>(out)[cell[u]] = o; }
    } else {
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) {
                W4 v = w4_zero(); v.w[0] = ch[u];
                if (MONT) v = fr_redc<1>(ch[u], 0, 0);
                store_cell_k<MONT>(out, cell[u], v, CK_RAW);
            }
    }
}

__device__ __forceinline__ void store_cell(u64* out, size_t cell, W4 v, int mont) {
    if (mont) v = fr_to_mont(v);
    ulonglong4 o; o.x = v.w[0]; o.y = v.w[1]; o.z = v.w[2]; o.w = v.w[3];
    reinterpret_cast<ulonglong4*>(out)[cell] = o;
}

// cell `sub` of an assigned value with a range check (A.1): 0 = the value, 1..k = lookup chunks, k+1 = shifted top chunk
__device__ __forceinline__ W4 range_cell(const W4& v, int sub, int bits, int lookup, int k) {
    if (sub == 0) return v;
    if (sub <= k) return w4_low(w4_shr(v, (sub - 1) * lookup), lookup);
    return w4_shl(w4_low(w4_shr(v, (k - 1) * lookup), lookup), lookup - bits % lookup);
}

// ---- k_cells_mulmod ---------------------------------------------------------------------------------------
// One CTA works on one group at a time.  Shared memory: limbs of a, b, q, rem, n^2 (2 words each), the three
// column arrays (4 words each), carries and cs (2 words each).
__global__ void __launch_bounds__(128) k_cells_mulmod(CellLayout Y, const u64* __restrict__ consts /* n2 limbs, word_max, q_acc, mod_acc */,
                                                      const u64* __restrict__ a, const u64* __restrict__ b,
                                                      const u64* __restrict__ q, const u64* __restrict__ rem,
                                                      size_t count, int words, int mont, u64* __restrict__ out, int* flags) {
    extern __shared__ u64 sm[];
    const int L = Y.L, NC = 2 * L - 1, lb = Y.limb_bits;
    u64* s_a = sm;                 // [L][2]
    u64* s_b = s_a + 2 * L;
    u64* s_q = s_b + 2 * L;
    u64* s_r = s_q + 2 * L;
    u64* s_n = s_r + 2 * L;
    W4* s_ab = (W4*)(s_n + 2 * L); // [NC]
    W4* s_qn = s_ab + NC;
    W4* s_qp = s_qn + NC;
    u64* s_carry = (u64*)(s_qp + NC);   // [NC][2]  carry_{i+1}
    u64* s_cs = s_carry + 2 * NC;       // [NC][2]
    __shared__ int s_eq;
    const bool wide = lb > 64;
    const u64* c_n2 = consts;                  // [L][2]
    const u64* c_wmax = consts + 2 * L;        // 4 words
    const u64* c_qacc = c_wmax + 4;            // [NC][2]
    const u64* c_macc = c_qacc + 2 * NC;       // [NC][2]
    for (int i = threadIdx.x; i < 2 * L; i += blockDim.x) s_n[i] = c_n2[i];
    for (size_t g = blockIdx.x; g < count; g += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
            u128c va = bits_of(a + g * words, words, i * lb, lb), vb = bits_of(b + g * words, words, i * lb, lb);
            u128c vq = bits_of(q + g * words, words, i * lb, lb), vr = bits_of(rem + g * words, words, i * lb, lb);
            s_a[2 * i] = (u64)va; s_a[2 * i + 1] = (u64)(va >> 64);
            s_b[2 * i] = (u64)vb; s_b[2 * i + 1] = (u64)(vb >> 64);
            s_q[2 * i] = (u64)vq; s_q[2 * i + 1] = (u64)(vq >> 64);
            s_r[2 * i] = (u64)vr; s_r[2 * i + 1] = (u64)(vr >> 64);
        }
        __syncthreads();
        // columns: thread i and thread NC-1-i together hold L+1 products per array -> pair the long with the short
        for (int i = threadIdx.x; i < NC; i += blockDim.x) {
            W4 ab = w4_zero(), qn = w4_zero();
            const int j0 = i < L ? 0 : i - L + 1, j1 = i < L ? i : L - 1;
            for (int j = j0; j <= j1; j++) {
                w4_mac(ab, s_a[2 * j], s_a[2 * j + 1], s_b[2 * (i - j)], s_b[2 * (i - j) + 1], wide);
                w4_mac(qn, s_q[2 * j], s_q[2 * j + 1], s_n[2 * (i - j)], s_n[2 * (i - j) + 1], wide);
            }
            s_ab[i] = ab; s_qn[i] = qn;
            if (i < L) { W4 r4; r4.w[0] = s_r[2 * i]; r4.w[1] = s_r[2 * i + 1]; r4.w[2] = r4.w[3] = 0; w4_add(qn, r4); }
            s_qp[i] = qn;
        }
        __syncthreads();
        // is_equal_muled carry chain (A.6): sequential over the columns
        if (threadIdx.x == 0) {
            W4 carry = w4_zero(), wmax; wmax.w[0] = c_wmax[0]; wmax.w[1] = c_wmax[1]; wmax.w[2] = c_wmax[2]; wmax.w[3] = c_wmax[3];
            int eq = 1, bad = 0;
            for (int i = 0; i < NC; i++) {
                W4 s = s_ab[i];
                w4_sub(s, s_qp[i]); w4_add(s, carry); w4_add(s, wmax);
                if (s.w[3] >> 63) { bad = 1; s = w4_zero(); }          // negative sum: the constraint would fail
                carry = w4_shr(s, lb);
                W4 cs = w4_low(s, lb);
                s_carry[2 * i] = carry.w[0]; s_carry[2 * i + 1] = carry.w[1];
                s_cs[2 * i] = cs.w[0]; s_cs[2 * i + 1] = cs.w[1];
                if (carry.w[2] | carry.w[3]) bad = 1;
                eq &= (cs.w[0] == c_macc[2 * i] && cs.w[1] == c_macc[2 * i + 1]);
                if (i == NC - 1) eq &= (carry.w[0] == c_qacc[2 * i] && carry.w[1] == c_qacc[2 * i + 1]);
            }
            s_eq = eq;
            if (bad || !eq) atomicOr(flags, 2);
        }
        __syncthreads();
        // cells, coalesced: thread c writes cell c of the group
        const size_t base = g * (size_t)Y.n_cells;
        for (int c = threadIdx.x; c < Y.n_cells; c += blockDim.x) {
            W4 v = w4_zero();
            if (c < Y.off_ab) {                                        // q limbs, then rem limbs, with range checks
                const int cc = c < Y.off_rem ? c : c - Y.off_rem;
                const u64* src = c < Y.off_rem ? s_q : s_r;
                const int limb = cc / Y.cpl, sub = cc - limb * Y.cpl;
                W4 x; x.w[0] = src[2 * limb]; x.w[1] = src[2 * limb + 1]; x.w[2] = x.w[3] = 0;
                v = range_cell(x, sub, lb, Y.lookup_bits, Y.kl);
            } else if (c < Y.off_qn) v = s_ab[c - Y.off_ab];
            else if (c < Y.off_qnp) v = s_qn[c - Y.off_qn];
            else if (c < Y.off_eq) v = s_qp[c - Y.off_qnp];
            else if (c == Y.n_cells - 1) v.w[0] = (u64)s_eq;
            else {
                const int cc = c - Y.off_eq;
                const int i = cc / Y.eq_stride, sub = cc - i * Y.eq_stride;
                if (sub == 0) { v.w[0] = s_carry[2 * i]; v.w[1] = s_carry[2 * i + 1]; }
                else if (sub == 1) { v.w[0] = s_cs[2 * i]; v.w[1] = s_cs[2 * i + 1]; }
                else if (sub == 2) { v.w[0] = c_qacc[2 * i]; v.w[1] = c_qacc[2 * i + 1]; }
                else if (sub == 3) { v.w[0] = c_macc[2 * i]; v.w[1] = c_macc[2 * i + 1]; }
                else {
                    W4 x = w4_zero(); x.w[0] = s_carry[2 * i]; x.w[1] = s_carry[2 * i + 1];
                    v = range_cell(x, sub - 3, Y.carry_bits, Y.lookup_bits, Y.kc);
                }
            }
            store_cell(out, base + c, v, mont);
        }
    }
}


// ---- k_cells_mulmod64: the production case, limb_bits = 64 --------------------------------------------------------
// One WARP per group, four independent warps per CTA (no CTA-wide barrier inside the loop, so the sequential carry chain
// of one group overlaps the column products and the stores of the others).  Per warp 4L + 3(2L-1) words of shared memory:
// the limbs of a, b, q, rem and d_i = ab_i - (qn_i + rem_i) + word_max, overwritten in place by (cs_i, carry_(i+1)).
// Columns: lane l owns the pair (c, c+L), c = l, l+32, ..: for j = 0..L-1 the product a_j * b_((c-j) mod L) belongs to
// column c when j <= c and to column c+L otherwise — every lane runs exactly L uniform iterations.
// All stores are one 32-byte cell per lane at consecutive addresses.
__device__ __forceinline__ void mac3p(u64& c0, u64& c1, u64& c2, u64 lo, u64 hi) {
    asm("add.cc.u64 %0, %0, %3; addc.cc.u64 %1, %1, %4; addc.u64 %2, %2, 0;" : "+l"(c0), "+l"(c1), "+l"(c2) : "l"(lo), "l"(hi));
}
__device__ __forceinline__ W4 chunk_cell(u64 lo, u64 hi, int sub, int pad, int lookup, int k) {    // sub >= 1; pad = lookup - bits % lookup
    const int idx = sub <= k ? sub - 1 : k - 1;
    const int sft = idx * lookup;
    const u64 x = sft >= 64 ? hi >> (sft - 64) : (sft ? (lo >> sft) | (hi << (64 - sft)) : lo);
    u64 ch = x & ((1ull << lookup) - 1);
    ch <<= (sub > k ? pad : 0);
    W4 r = w4_zero(); r.w[0] = ch;
    return r;
}

// 192-bit accumulator of 64 x 64 -> 128-bit products on 32-bit words, carry pairs that ptxas fuses into IMAD.WIDE.U32(.X):
// products at even word offsets (a0 b0 at word 0, a1 b1 at word 2) go to e0..e4, the two cross products (word 1) to o1..o3, kept
// apart so that every multiply-add lands on an aligned register pair; value = E + (O << 32).  Up to 2^31 products.
struct Acc192 {
    unsigned e0, e1, e2, e3, e4, o1, o2, o3;
    __device__ __forceinline__ void clear() { e0 = e1 = e2 = e3 = e4 = o1 = o2 = o3 = 0; }
    __device__ __forceinline__ void words(u64& w0, u64& w1, u64& w2) const {
        w0 = ((u64)e1 << 32) | e0; w1 = ((u64)e3 << 32) | e2; w2 = e4;
        const u64 x0 = (u64)o1 << 32, x1 = ((u64)o3 << 32) | o2;
        asm("add.cc.u64 %0, %0, %3; addc.cc.u64 %1, %1, %4; addc.u64 %2, %2, 0;" : "+l"(w0), "+l"(w1), "+l"(w2) : "l"(x0), "l"(x1));
    }
};
__device__ __forceinline__ void mac64(Acc192& A, u64 a, u64 b) {
    const unsigned a0 = (unsigned)a, a1 = (unsigned)(a >> 32), b0 = (unsigned)b, b1 = (unsigned)(b >> 32);
    asm("mad.lo.cc.u32 %0, %5, %7, %0; madc.hi.cc.u32 %1, %5, %7, %1; madc.lo.cc.u32 %2, %6, %8, %2; madc.hi.cc.u32 %3, %6, %8, %3; addc.u32 %4, %4, 0;"
        : "+r"(A.e0), "+r"(A.e1), "+r"(A.e2), "+r"(A.e3), "+r"(A.e4) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" : "+r"(A.o1), "+r"(A.o2), "+r"(A.o3) : "r"(a0), "r"(b1));
    asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" : "+r"(A.o1), "+r"(A.o2), "+r"(A.o3) : "r"(a1), "r"(b0));
}

#ifndef PB200_CELLS_MONT_OCC
#define PB200_CELLS_MONT_OCC 5      // resident CTAs per SM the Montgomery variant is compiled for (A/B: tools/build_variant.py)
#endif
template <bool MONT>
__global__ void __launch_bounds__(128, MONT ? PB200_CELLS_MONT_OCC : 6) k_cells_mulmod64(CellLayout Y, u64 m_cpl, u64 m_eq, u64 m_eqs, u64 m_ch, const u64* __restrict__ consts,
                                                           const u64* __restrict__ a, const u64* __restrict__ b,
                                                           const u64* __restrict__ q, const u64* __restrict__ rem,
                                                           size_t count, u64* __restrict__ out, int* flags, const u64* __restrict__ mtab) {
    extern __shared__ u64 sm[];
    const int L = Y.L, NC = 2 * L - 1;
    const u64* c_accm = consts + 2 * L + 4 + 4 * NC;      // Montgomery forms of q_acc, mod_acc: [NC][4] each
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* s_n = sm;
    // per warp: two limb sets (double buffer), then d
    u64* s_w = sm + L + (size_t)warp * (8 * L + 3 * NC + 1);
    u64* s_d = s_w + 8 * L;
    const u64* c_wmax = consts + 2 * L;        // consts: n2 limbs [L][2], word_max (4 words), q_acc [NC][2], mod_acc [NC][2]
    const u64* c_qacc = c_wmax + 4;
    const u64* c_macc = c_qacc + 2 * NC;
    for (int i = threadIdx.x; i < L; i += blockDim.x) s_n[i] = consts[2 * i];
    __syncthreads();
    const u64 wm0 = c_wmax[0], wm1 = c_wmax[1], wm2 = c_wmax[2];
    const int pad_l = Y.lookup_bits ? Y.lookup_bits - 64 % Y.lookup_bits : 0, pad_c = Y.lookup_bits ? Y.lookup_bits - Y.carry_bits % Y.lookup_bits : 0;
    // the limbs of the NEXT group travel global -> shared with cp.async while the current group is expanded
    auto prefetch = [&](size_t g, int buf) {
        u64* dst = s_w + buf * 4 * L;
        for (int i = lane; i < L; i += 32) {
            const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d0), "l"(a + g * L + i));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d0 + 8 * L), "l"(b + g * L + i));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d0 + 16 * L), "l"(q + g * L + i));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d0 + 24 * L), "l"(rem + g * L + i));
        }
        asm volatile("cp.async.commit_group;");
    };
    const size_t g_first = (size_t)blockIdx.x * 4 + warp, g_step = (size_t)gridDim.x * 4;
    int buf = 0;
    if (g_first < count) prefetch(g_first, 0);
    for (size_t g = g_first; g < count; g += g_step, buf ^= 1) {
        asm volatile("cp.async.wait_group 0;");
        __syncwarp();
        if (g + g_step < count) prefetch(g + g_step, buf ^ 1);
        u64* s_a = s_w + buf * 4 * L;
        u64* s_b = s_a + L; u64* s_q = s_b + L; u64* s_r = s_q + L;
        const size_t base = g * (size_t)Y.n_cells;
        // q and rem: limb cells with their range-check chunks
        if (!MONT) {
            for (int c = lane; c < Y.off_ab; c += 32) {
                const int cc = c < Y.off_rem ? c : c - Y.off_rem;
                const u64* src = c < Y.off_rem ? s_q : s_r;
                const int limb = (int)(((u64)cc * m_cpl) >> 32), sub = cc - limb * Y.cpl;
                const u64 x = src[limb];
                if (sub == 0) { W4 v = w4_zero(); v.w[0] = x; store_cell_k<false>(out, base + c, v, CK_RAW); }
                else store_chunk<false>(out, base + c, chunk_cell(x, 0, sub, pad_l, Y.lookup_bits, Y.kl), mtab);
            }
        } else {
            // the limb cells themselves: one-word conversions, all lanes on the same path
            for (int c = lane; c < 2 * L; c += 32) {
                const int limb = c < L ? c : c - L;
                store_cell_k<true>(out, base + (c < L ? 0 : Y.off_rem) + (size_t)limb * Y.cpl, fr_redc<1>((c < L ? s_q : s_r)[limb], 0, 0), CK_RAW);
            }
            // their range-check chunks: table gathers, four per lane and pass
            const int cpk = Y.cpl - 1, n_ch = 2 * L * cpk;
            for (int i0 = lane; i0 < n_ch; i0 += 128) {
                size_t cell[4]; u64 ch[4]; bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int idx = i0 + 32 * u;
                    ok[u] = idx < n_ch;
                    const int li = ok[u] ? (int)(((u64)idx * m_ch) >> 32) : 0, sub = ok[u] ? idx - li * cpk + 1 : 1;      // li: limb index over q then rem
                    const u64 x = li < L ? s_q[li] : s_r[li - L];
                    ch[u] = chunk_cell(x, 0, sub, pad_l, Y.lookup_bits, Y.kl).w[0];
                    cell[u] = base + (li < L ? 0 : Y.off_rem) + (size_t)(li < L ? li : li - L) * Y.cpl + sub;
                }
                store_chunks4<true>(out, cell, ch, ok, mtab);
            }
        }
        // no-carry columns ab and q*n^2, the sums, and d
        for (int c = lane; c < L; c += 32) {
            // one running accumulator per product; its value when j reaches c + 1 is column c, the rest is column c + L.
            // The lanes of a warp switch at j = c + 1 in [k0 + 1, k0 + 32] (k0 = c - lane is warp-uniform): only that segment of
            // the loop carries the snapshot.
            const int k0 = c - lane;
            Acc192 A, Q, As, Qs;
            A.clear(); Q.clear(); As.clear(); Qs.clear();
            int bi = c, j = 0;
#define PB200_COL_STEP() { const u64 aj = s_a[j], qj = s_q[j], bv = s_b[bi], nv = s_n[bi]; mac64(A, aj, bv); mac64(Q, qj, nv); bi = bi ? bi - 1 : L - 1; }
#pragma unroll 4
            for (; j <= k0; j++) PB200_COL_STEP()
            const int j_sw = min(k0 + 32, L - 1);
#pragma unroll 2
            for (; j <= j_sw; j++) { if (j == c + 1) { As = A; Qs = Q; } PB200_COL_STEP() }
#pragma unroll 4
            for (; j < L; j++) PB200_COL_STEP()
#undef PB200_COL_STEP
            if (c == L - 1) { As = A; Qs = Q; }
            u64 ah0, ah1, ah2, qh0, qh1, qh2, al0, al1, al2, ql0, ql1, ql2;
            A.words(ah0, ah1, ah2); Q.words(qh0, qh1, qh2); As.words(al0, al1, al2); Qs.words(ql0, ql1, ql2);
            asm("sub.cc.u64 %0, %0, %3; subc.cc.u64 %1, %1, %4; subc.u64 %2, %2, %5;" : "+l"(ah0), "+l"(ah1), "+l"(ah2) : "l"(al0), "l"(al1), "l"(al2));
            asm("sub.cc.u64 %0, %0, %3; subc.cc.u64 %1, %1, %4; subc.u64 %2, %2, %5;" : "+l"(qh0), "+l"(qh1), "+l"(qh2) : "l"(ql0), "l"(ql1), "l"(ql2));
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int col = c + h * L;
                if (col < NC) {
                    W4 ab, qn;
                    ab.w[0] = h ? ah0 : al0; ab.w[1] = h ? ah1 : al1; ab.w[2] = h ? ah2 : al2; ab.w[3] = 0;
                    qn.w[0] = h ? qh0 : ql0; qn.w[1] = h ? qh1 : ql1; qn.w[2] = h ? qh2 : ql2; qn.w[3] = 0;
                    W4 qp = qn;
                    if (h == 0) mac3p(qp.w[0], qp.w[1], qp.w[2], s_r[c], 0);
                    // (Montgomery output: the canonical columns are parked in their own output cells — 8 KB per group that stay in L2 —
                    // and converted in place by the passes below; keeping them in shared memory cost a third of the resident warps)
                    store_cell_k<false>(out, base + Y.off_ab + col, ab, CK_RAW);
                    store_cell_k<false>(out, base + Y.off_qn + col, qn, CK_RAW);
                    if (!MONT) store_cell_k<false>(out, base + Y.off_qnp + col, qp, CK_RAW);
                    // d = ab - qp + word_max (two's complement over 3 words)
                    u64 d0, d1, d2;
                    asm("sub.cc.u64 %0, %3, %6; subc.cc.u64 %1, %4, %7; subc.u64 %2, %5, %8;"
                        : "=l"(d0), "=l"(d1), "=l"(d2) : "l"(ab.w[0]), "l"(ab.w[1]), "l"(ab.w[2]), "l"(qp.w[0]), "l"(qp.w[1]), "l"(qp.w[2]));
                    asm("add.cc.u64 %0, %0, %3; addc.cc.u64 %1, %1, %4; addc.u64 %2, %2, %5;" : "+l"(d0), "+l"(d1), "+l"(d2) : "l"(wm0), "l"(wm1), "l"(wm2));
                    s_d[3 * col] = d0; s_d[3 * col + 1] = d1; s_d[3 * col + 2] = d2;
                }
            }
        }
        __syncwarp();
        // is_equal_muled carry chain (A.6): s_i = d_i + carry_i; cs_i = s_i mod 2^64, carry_(i+1) = s_i >> 64, in place
        int eq = 1;
        if (lane == 0) {
            u64 c0 = 0, c1 = 0;
            int bad = 0;
            for (int i = 0; i < NC; i++) {
                u64 s0 = s_d[3 * i], s1 = s_d[3 * i + 1], s2 = s_d[3 * i + 2];
                asm("add.cc.u64 %0, %0, %3; addc.cc.u64 %1, %1, %4; addc.u64 %2, %2, 0;" : "+l"(s0), "+l"(s1), "+l"(s2) : "l"(c0), "l"(c1));
                bad |= (int)(s2 >> 63);
                c0 = s1; c1 = s2;
                s_d[3 * i + 1] = s1; s_d[3 * i + 2] = s2; s_d[3 * i] = s0;
            }
            eq &= (c0 == c_qacc[2 * (NC - 1)] && c1 == c_qacc[2 * (NC - 1) + 1]);
            if (bad || !eq) atomicOr(flags, 2);
        }
        eq = __shfl_sync(0xffffffffu, eq, 0);
        __syncwarp();
        {
            int ok = 1;
            for (int i = lane; i < NC; i += 32) ok &= (s_d[3 * i] == __ldg(c_macc + 2 * i));
            ok = __all_sync(0xffffffffu, ok);
            if (!ok && eq && lane == 0) atomicOr(flags, 2);
            eq &= ok;
        }
        if (!MONT) {
            // canonical output: one loop over the eq section's cells, lane c writes cell c (contiguous 1 KB per warp store)
            const int n_eq = Y.n_cells - Y.off_eq;
            for (int cc = lane; cc < n_eq; cc += 32) {
                W4 v = w4_zero();
                if (cc == n_eq - 1) v.w[0] = (u64)eq;
                else {
                    const int i = (int)(((u64)cc * m_eqs) >> 32), sub = cc - i * Y.eq_stride;
                    const u64 cs = s_d[3 * i], k0 = s_d[3 * i + 1], k1 = s_d[3 * i + 2];
                    if (sub >= 4) v = chunk_cell(k0, k1, sub - 3, pad_c, Y.lookup_bits, Y.kc);
                    else if (sub >= 2) {
                        const u64* src = sub == 2 ? c_qacc : c_macc;
                        v.w[0] = __ldg(src + 2 * i); v.w[1] = __ldg(src + 2 * i + 1);
                    } else { v.w[0] = sub ? cs : k0; v.w[1] = sub ? 0 : k1; }
                }
                store_cell_k<false>(out, base + Y.off_eq + cc, v, CK_RAW);
            }
            __syncwarp();
            continue;
        }
        // Montgomery output: conversion passes of uniform cell kind (no divergence between the conversion paths, one call site each).
        // Columns: Montgomery form is linear, so the cell of qn + rem is the cell of qn plus the (one-word) cell of rem, and for the
        // upper columns it IS the cell of qn: one general conversion serves two cells.
        __syncwarp();        // (small L: the rem limb cells read back below were written by other lanes of this warp)
        for (int col = lane; col < NC; col += 32) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(out) + 2 * (base + Y.off_ab + col);    // written by this lane above
            const ulonglong2 x01 = __ldcg(src), x23 = __ldcg(src + 1);
            store_cell_k<true>(out, base + Y.off_ab + col, fr_redc<3>(x01.x, x01.y, x23.x), CK_RAW);
        }
        for (int col = lane; col < NC; col += 32) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(out) + 2 * (base + Y.off_qn + col);
            const ulonglong2 x01 = __ldcg(src), x23 = __ldcg(src + 1);
            const W4 qn_m = fr_redc<3>(x01.x, x01.y, x23.x);
            store_cell_k<true>(out, base + Y.off_qn + col, qn_m, CK_RAW);
            W4 qp_m = qn_m;
            if (col < L) {       // + the cell of rem_col, which the limb pass above already converted and wrote (L2 hit)
                const ulonglong2* rc = reinterpret_cast<const ulonglong2*>(out) + 2 * (base + Y.off_rem + (size_t)col * Y.cpl);
                const ulonglong2 r01 = __ldcg(rc), r23 = __ldcg(rc + 1);
                W4 rm; rm.w[0] = r01.x; rm.w[1] = r01.y; rm.w[2] = r23.x; rm.w[3] = r23.y;
                qp_m = fr_add(qn_m, rm);
            }
            store_cell_k<true>(out, base + Y.off_qnp + col, qp_m, CK_RAW);
        }
        // eq section, per column [carry (two words), cs (one word), q_acc, mod_acc (per-key, pre-converted)], then the carries' chunks
        for (int i = lane; i < NC; i += 32) {
            const size_t cell = base + Y.off_eq + (size_t)i * Y.eq_stride;
            store_cell_k<true>(out, cell, fr_redc<2>(s_d[3 * i + 1], s_d[3 * i + 2], 0), CK_RAW);
            W4 acc_m[2];
#pragma unroll
            for (int w = 0; w < 2; w++) {
                const u64* m4 = c_accm + 4 * ((size_t)w * NC + i);
                acc_m[w].w[0] = __ldg(m4); acc_m[w].w[1] = __ldg(m4 + 1); acc_m[w].w[2] = __ldg(m4 + 2); acc_m[w].w[3] = __ldg(m4 + 3);
            }
            // cs_i equals mod_acc_i whenever the chip's equality holds (checked above): its cell is then the pre-converted constant
            const bool same = s_d[3 * i] == __ldg(c_macc + 2 * i);
            store_cell_k<true>(out, cell + 1, same ? acc_m[1] : fr_redc<1>(s_d[3 * i], 0, 0), CK_RAW);
            store_cell_k<true>(out, cell + 2, acc_m[0], CK_RAW);
            store_cell_k<true>(out, cell + 3, acc_m[1], CK_RAW);
        }
        const int cpc = Y.kc + Y.xc, n_cc = (NC - 1) * cpc;
        for (int i0 = lane; i0 < n_cc; i0 += 128) {
            size_t cell[4]; u64 ch[4]; bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = i0 + 32 * u;
                ok[u] = idx < n_cc;
                const int i = ok[u] ? (int)(((u64)idx * m_eq) >> 32) : 0, sub = ok[u] ? idx - i * cpc : 0;     // m_eq: magic of cpc
                ch[u] = chunk_cell(s_d[3 * i + 1], s_d[3 * i + 2], sub + 1, pad_c, Y.lookup_bits, Y.kc).w[0];
                cell[u] = base + Y.off_eq + (size_t)i * Y.eq_stride + 4 + sub;
            }
            store_chunks4<true>(out, cell, ch, ok, mtab);
        }
        if (lane == 0) {       // the flag cell: Montgomery form of 1 is R mod p
            W4 v = w4_zero();
            if (eq) { v.w[0] = FR_R1[0]; v.w[1] = FR_R1[1]; v.w[2] = FR_R1[2]; v.w[3] = FR_R1[3]; }
            store_cell_k<true>(out, base + Y.n_cells - 1, v, CK_RAW);
        }
        __syncwarp();
    }
}

// ---- k_cells_assign: assign_integer of `count` values ---------------------------------------------------------
__global__ void k_cells_assign(const u64* __restrict__ vals, size_t count, int words, int nl, int limb_bits, int lookup, int k,
                               int cpl, int mont, u64* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = (size_t)nl * cpl;
    if (idx >= count * per) return;
    const size_t u = idx / per; const int cc = (int)(idx - u * per);
    const int limb = cc / cpl, sub = cc - limb * cpl;
    W4 x = w4_from128(bits_of(vals + u * words, words, limb * limb_bits, limb_bits));
    store_cell(out, idx, range_cell(x, sub, limb_bits, lookup, k), mont);
}

// ---- k_cells_n2: square(n) + refresh, one thread (per key) ------------------------------------------------------
__global__ void k_cells_n2(const u64* __restrict__ n_words, int words, int kn, int limb_bits, int lookup, int kl, int xl,
                           const int* __restrict__ inc, int n_out, int mont, u64* __restrict__ out, int* n_written, int* flags) {
    if (blockIdx.x || threadIdx.x) return;
    extern __shared__ u64 smx[];
    W4* x = (W4*)smx;                       // n_out entries
    const int NC = 2 * kn - 1;
    const bool wide = limb_bits > 64;
    size_t cell = 0;
    for (int i = 0; i < n_out; i++) x[i] = w4_zero();
    for (int i = 0; i < NC; i++) {          // square: no-carry columns (A.2)
        W4 acc = w4_zero();
        const int j0 = i < kn ? 0 : i - kn + 1, j1 = i < kn ? i : kn - 1;
        for (int j = j0; j <= j1; j++) {
            u128c aj = bits_of(n_words, words, j * limb_bits, limb_bits), bj = bits_of(n_words, words, (i - j) * limb_bits, limb_bits);
            w4_mac(acc, (u64)aj, (u64)(aj >> 64), (u64)bj, (u64)(bj >> 64), wide);
        }
        x[i] = acc;
        store_cell(out, cell++, acc, mont);
    }
    for (int i = 0; i < NC; i++) {          // refresh (A.3)
        W4 limb = x[i];
        for (int j = 0; j <= inc[i]; j++) {
            W4 qq = w4_shr(limb, limb_bits), rr = w4_low(limb, limb_bits);
            store_cell(out, cell++, qq, mont);
            store_cell(out, cell++, rr, mont);
            if (j == 0) x[i] = rr; else if (i + j < n_out) w4_add(x[i + j], rr);
            limb = qq;
        }
        if (limb.w[0] | limb.w[1] | limb.w[2] | limb.w[3]) atomicOr(flags, 2);
    }
    if (lookup) {
        for (int i = 0; i < n_out; i++)
            for (int sub = 1; sub <= kl + xl; sub++) store_cell(out, cell++, range_cell(x[i], sub, limb_bits, lookup, kl), mont);
    }
    *n_written = (int)cell;
}

// ---- launchers ----------------------------------------------------------------------------------------------------
size_t cells_mulmod_smem(const CellLayout& Y) {
    const size_t L = Y.L, NC = 2 * L - 1;
    return (5 * 2 * L + 3 * 4 * NC + 2 * 2 * NC) * sizeof(u64);
}
cudaError_t cells_mont_table(int lookup_bits, u64* d_tab, cudaStream_t st) {
    const int n = 1 << lookup_bits;
    k_mont_table<<<(n + 255) / 256, 256, 0, st>>>(n, d_tab);
    count_launch();
    return cudaGetLastError();
}
cudaError_t cells_mulmod(const CellLayout& Y, const u64* d_consts, const u64* d_a, const u64* d_b, const u64* d_q, const u64* d_rem,
                         size_t count, int words, int mont, u64* d_out, int* d_flags, int sms, const u64* d_mtab, cudaStream_t st) {
    if (!count) return cudaSuccess;
    if (Y.limb_bits == 64 && Y.n_cells < 65536) {
        const size_t L = Y.L, NC = 2 * L - 1;
        const size_t smem64 = (L + 4 * (8 * L + 3 * NC + 1)) * sizeof(u64);
        {   // per device (a function attribute belongs to the current context): set on every launch, it is a host-side table write
            cudaError_t e = mont ? cudaFuncSetAttribute(k_cells_mulmod64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)
                                 : cudaFuncSetAttribute(k_cells_mulmod64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            if (e != cudaSuccess) return e;
        }
        const u64 m_eqs = (0x100000000ull + Y.eq_stride - 1) / Y.eq_stride;
        const u64 m_cpl = (0x100000000ull + Y.cpl - 1) / Y.cpl, m_eq = (Y.kc + Y.xc) ? (0x100000000ull + (Y.kc + Y.xc) - 1) / (Y.kc + Y.xc) : 0;
        const u64 m_ch = Y.cpl > 1 ? (0x100000000ull + (Y.cpl - 1) - 1) / (Y.cpl - 1) : 0;
        size_t ctas = (count + 3) / 4;
        const size_t per_sm = smem64 ? (200 * 1024) / smem64 : 6;
        const size_t occ = mont ? PB200_CELLS_MONT_OCC : 6;
        const size_t cap = (size_t)sms * (per_sm < 1 ? 1 : (per_sm > occ ? occ : per_sm));
        if (ctas > cap) ctas = cap;
        if (mont) k_cells_mulmod64<true><<<(unsigned)ctas, 128, smem64, st>>>(Y, m_cpl, m_eq, m_eqs, m_ch, d_consts, d_a, d_b, d_q, d_rem, count, d_out, d_flags, d_mtab);
        else k_cells_mulmod64<false><<<(unsigned)ctas, 128, smem64, st>>>(Y, m_cpl, m_eq, m_eqs, m_ch, d_consts, d_a, d_b, d_q, d_rem, count, d_out, d_flags, nullptr);
        count_launch();
        return cudaGetLastError();
    }
    const size_t smem = cells_mulmod_smem(Y);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_cells_mulmod, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
    }
    size_t grid = count < (size_t)sms * 8 ? count : (size_t)sms * 8;
    k_cells_mulmod<<<(unsigned)grid, 128, smem, st>>>(Y, d_consts, d_a, d_b, d_q, d_rem, count, words, mont, d_out, d_flags);
    count_launch();
    return cudaGetLastError();
}
cudaError_t cells_assign(const u64* d_vals, size_t count, int words, int nl, int limb_bits, int lookup, int k, int cpl, int mont,
                         u64* d_out, cudaStream_t st) {
    const size_t total = count * (size_t)nl * cpl;
    if (!total) return cudaSuccess;
    k_cells_assign<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_vals, count, words, nl, limb_bits, lookup, k, cpl, mont, d_out);
    count_launch();
    return cudaGetLastError();
}
cudaError_t cells_n2(const u64* d_n_words, int words, int kn, int limb_bits, int lookup, int kl, int xl, const int* d_inc, int n_out,
                     int mont, u64* d_out, int* d_n_written, int* d_flags, cudaStream_t st) {
    k_cells_n2<<<1, 1, (size_t)n_out * sizeof(W4), st>>>(d_n_words, words, kn, limb_bits, lookup, kl, xl, d_inc, n_out, mont, d_out,
                                                        d_n_written, d_flags);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pb200
