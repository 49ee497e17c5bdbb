// simple64.cuh — "simple64" engine: one ciphertext per thread, 64-bit limbs in local memory.
//
// This is the always-available CUDA path: any key size up to PB200_SIMPLE_MAXK words, exact
// (q, rem) for every mul_mod, and the reference's own LSB-first chain
// (BigUintChip::pow_mod_fixed_exp; driven from src/paillier.rs:51,55,57).
// It is not the fast path (see block28.cuh); it is the witness producer for small batches, the
// engine for sizes the block engine is not compiled for, and the on-GPU cross-check of block28.
//
// mul_mod(a, b) by Barrett reduction against the normalised modulus Nt = n^2 << s (top bit of the
// k-word container set; HAC 14.42 with b = 2^64):
//     X  = a*b            (2k words)          Xs = X << s
//     q1 = Xs >> 64(k-1)  (k+1 words)         q3 = (q1 * mu) >> 64(k+1),  mu = floor(2^(128k)/Nt)
//     r  = Xs - q3*Nt  mod 2^(64(k+1));       while r >= Nt: r -= Nt, q3 += 1        (<= 2 times)
//     q  = q3 = floor(a*b / n^2),             rem = r >> s = a*b mod n^2
// because floor(Xs/Nt) = floor(X/n^2) and Xs mod Nt = (X mod n^2) << s.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define PB200_SIMPLE_MAXK 128   // words of the container: |n| up to 4096 bits
#define PB200_DIGEST_INIT 0xcbf29ce484222325ull
#define PB200_DIGEST_PRIME 0x100000001b3ull
#define PB200_DIGEST_C 0x9E3779B97F4A7C15ull

namespace pb200 {

typedef uint64_t u64;
typedef unsigned __int128 u128;

struct SimpleConsts {
    int k;          // words of the container (= words_out)
    int kin;        // words of n, g, m, r (= words_in)
    int s;          // normalisation shift: Nt = n2 << s has bit 64k-1 set
    int n_bits;     // enc_bits
    int exp_bits;   // bits(n): length of the r-chain
    u64 Nt[PB200_SIMPLE_MAXK];
    u64 mu[PB200_SIMPLE_MAXK + 1];
    u64 n[PB200_SIMPLE_MAXK / 2];   // exponent of the r-chain
    u64 g[PB200_SIMPLE_MAXK / 2];   // base of the g-chain (as assigned, not reduced)
};

// (c0,c1,c2) += a*b
__device__ __forceinline__ void mac3(u64& c0, u64& c1, u64& c2, u64 a, u64 b) {
    u64 lo = a * b, hi = __umul64hi(a, b);
    asm("add.cc.u64 %0, %0, %3; addc.cc.u64 %1, %1, %4; addc.u64 %2, %2, 0;"
        : "+l"(c0), "+l"(c1), "+l"(c2) : "l"(lo), "l"(hi));
}

// out[0..na+nb) = a[0..na) * b[0..nb)   (Comba, columns low to high)
static __device__ __noinline__ void mul_full(u64* out, const u64* a, int na, const u64* b, int nb) {
    u64 c0 = 0, c1 = 0, c2 = 0;
    for (int col = 0; col < na + nb - 1; col++) {
        int i0 = col - (nb - 1); if (i0 < 0) i0 = 0;
        int i1 = col < na - 1 ? col : na - 1;
        for (int i = i0; i <= i1; i++) mac3(c0, c1, c2, a[i], b[col - i]);
        out[col] = c0; c0 = c1; c1 = c2; c2 = 0;
    }
    out[na + nb - 1] = c0;
}

// out[0..nout) = columns [lo_col, lo_col+nout) of a*b, exact (carries from lower columns included)
static __device__ __noinline__ void mul_cols(u64* out, int lo_col, int nout, const u64* a, int na, const u64* b, int nb) {
    u64 c0 = 0, c1 = 0, c2 = 0;
    int last = lo_col + nout;  // exclusive
    for (int col = 0; col < last; col++) {
        if (col < na + nb - 1) {
            int i0 = col - (nb - 1); if (i0 < 0) i0 = 0;
            int i1 = col < na - 1 ? col : na - 1;
            for (int i = i0; i <= i1; i++) mac3(c0, c1, c2, a[i], b[col - i]);
        }
        if (col >= lo_col) out[col - lo_col] = c0;
        c0 = c1; c1 = c2; c2 = 0;
    }
}

__device__ __forceinline__ void shl_words(u64* x, int n, int s) {  // x <<= s, 0 <= s < 64n, in place
    int ws = s >> 6, bs = s & 63;
    for (int i = n - 1; i >= 0; i--) {
        u64 hi = i - ws >= 0 ? x[i - ws] : 0, lo = i - ws - 1 >= 0 ? x[i - ws - 1] : 0;
        x[i] = bs ? (hi << bs) | (lo >> (64 - bs)) : hi;
    }
}
__device__ __forceinline__ void shr_words(u64* out, const u64* x, int n, int s, int nout) {  // out = x >> s
    int ws = s >> 6, bs = s & 63;
    for (int i = 0; i < nout; i++) {
        u64 lo = i + ws < n ? x[i + ws] : 0, hi = i + ws + 1 < n ? x[i + ws + 1] : 0;
        out[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
}

// q[0..k), rem[0..k) <- floor(X / N), X mod N for the 2k-word X (X[2k] is scratch; X is destroyed), N the modulus the constants
// were built for.  Returns 1 if the quotient does not fit k words, else 0.
static __device__ __noinline__ int barrett_simple(const SimpleConsts* __restrict__ K, u64* q, u64* rem, u64* X) {
    const int k = K->k;
    u64 q3[PB200_SIMPLE_MAXK + 2];
    u64 r[PB200_SIMPLE_MAXK + 1];
    X[2 * k] = 0;
    // X << s must stay inside 2k words.  Operands that are not reduced (pb200_add_batch accepts them) can make X >= 2^(128k - s);
    // then q = floor(X / n^2) > 2^(128k - s) / 2^(64k - s) = 2^(64k) does not fit k words: exactly the range failure, reported
    // here before the shift would drop the top bits
    if (K->s) {
        const int top = 128 * k - K->s, w = top >> 6, b0 = top & 63;
        u64 over = X[w] >> b0;
        for (int i = w + 1; i < 2 * k; i++) over |= X[i];
        if (over) {
            for (int i = 0; i < k; i++) { q[i] = 0; rem[i] = 0; }
            return 1;
        }
    }
    shl_words(X, 2 * k, K->s);
    // q3 = (q1 * mu) >> 64(k+1); q1 = X[k-1 .. 2k] (k+1 words), mu: k+1 words
    mul_cols(q3, k + 1, k + 1, X + (k - 1), k + 1, K->mu, k + 1);
    // r = X[0..k] - (q3*Nt)[0..k]  mod 2^(64(k+1))
    mul_cols(r, 0, k + 1, q3, k + 1, K->Nt, k);
    {
        u64 borrow = 0;
        for (int i = 0; i <= k; i++) {
            u64 xi = X[i], pi = r[i];
            u64 d = xi - pi, b1 = xi < pi;
            u64 d2 = d - borrow, b2 = d < borrow;
            r[i] = d2; borrow = b1 | b2;
        }
    }
    // corrections (at most 2 by the Barrett bound; loop kept general)
    for (int it = 0; it < 4; it++) {
        bool ge = r[k] != 0;
        if (!ge) {
            ge = true;
            for (int i = k - 1; i >= 0; i--) {
                if (r[i] != K->Nt[i]) { ge = r[i] > K->Nt[i]; break; }
            }
        }
        if (!ge) break;
        u64 borrow = 0;
        for (int i = 0; i <= k; i++) {
            u64 ni = i < k ? K->Nt[i] : 0;
            u64 d = r[i] - ni, b1 = r[i] < ni;
            u64 d2 = d - borrow, b2 = d < borrow;
            r[i] = d2; borrow = b1 | b2;
        }
        for (int i = 0; i <= k; i++) { q3[i] += 1; if (q3[i] != 0) break; }
    }
    shr_words(rem, r, k + 1, K->s, k);
    for (int i = 0; i < k; i++) q[i] = q3[i];
    return q3[k] != 0;
}

// q[0..k), rem[0..k) <- floor(a*b/n2), a*b mod n2.  a, b: k words (zero-extended).  Returns 1 if
// the quotient does not fit k words (the reference's range check on q would fail), else 0.
static __device__ __noinline__ int mulmod_simple(const SimpleConsts* __restrict__ K, u64* q, u64* rem, const u64* a, const u64* b) {
    const int k = K->k;
    u64 X[2 * PB200_SIMPLE_MAXK + 1];
    mul_full(X, a, k, b, k);
    return barrett_simple(K, q, rem, X);
}

__device__ __forceinline__ u64 record_hash(const u64* q, const u64* rem, int k) {
    u64 h = 0, c = PB200_DIGEST_C;
    for (int i = 0; i < k; i++) { h += q[i] * c; c *= PB200_DIGEST_C; }
    for (int i = 0; i < k; i++) { h += rem[i] * c; c *= PB200_DIGEST_C; }
    return h;
}

// Sink for witness records of one unit: stores to global memory (rec != nullptr) and/or folds into the digest.
struct RecordSink {
    u64* rec;      // next record slot (2k words) or nullptr
    u64 digest;
    int k;
    __device__ __forceinline__ void emit(const u64* q, const u64* rem) {
        if (rec) {
            for (int i = 0; i < k; i++) rec[i] = q[i];
            for (int i = 0; i < k; i++) rec[k + i] = rem[i];
            rec += 2 * k;
        }
        digest = (digest ^ record_hash(q, rem, k)) * PB200_DIGEST_PRIME;
    }
};

}  // namespace pb200
