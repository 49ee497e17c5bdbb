// host_bigint.hpp — small exact unsigned big-integer type for per-key setup on the host.
//
// Per-key constants (n^2, the normalised modulus, Barrett reciprocals) are computed once per
// pb200_key on the CPU; everything per-ciphertext runs on the GPU.  The reference computes the same
// quantity n*n with num-bigint at src/paillier.rs:88,95 and in-circuit via square+refresh at :39-45.
// Little-endian u32 limbs; value semantics; no allocation tricks — this is not a hot path.
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>
#include <stdexcept>

namespace pb200 {

struct BigInt {
    std::vector<uint32_t> w;  // little-endian, normalised (no leading zero limbs)

    BigInt() {}
    explicit BigInt(uint64_t v) { if (v) { w.push_back((uint32_t)v); if (v >> 32) w.push_back((uint32_t)(v >> 32)); } }
    static BigInt from_u64_le(const uint64_t* p, size_t n) {
        BigInt r; r.w.resize(2 * n);
        for (size_t i = 0; i < n; i++) { r.w[2 * i] = (uint32_t)p[i]; r.w[2 * i + 1] = (uint32_t)(p[i] >> 32); }
        r.trim(); return r;
    }
    static BigInt from_u32_le(const uint32_t* p, size_t n) { BigInt r; r.w.assign(p, p + n); r.trim(); return r; }
    void to_u32_le(uint32_t* p, size_t n) const {
        if (w.size() > n) throw std::runtime_error("BigInt::to_u32_le overflow");
        for (size_t i = 0; i < n; i++) p[i] = i < w.size() ? w[i] : 0;
    }
    void to_u64_le(uint64_t* p, size_t n) const {
        if (w.size() > 2 * n) throw std::runtime_error("BigInt::to_u64_le overflow");
        for (size_t i = 0; i < n; i++) {
            uint64_t lo = 2 * i < w.size() ? w[2 * i] : 0, hi = 2 * i + 1 < w.size() ? w[2 * i + 1] : 0;
            p[i] = lo | (hi << 32);
        }
    }
    void trim() { while (!w.empty() && w.back() == 0) w.pop_back(); }
    bool is_zero() const { return w.empty(); }
    bool is_odd() const { return !w.empty() && (w[0] & 1); }
    size_t bits() const {
        if (w.empty()) return 0;
        uint32_t t = w.back(); size_t b = 0; while (t) { b++; t >>= 1; }
        return (w.size() - 1) * 32 + b;
    }
    bool bit(size_t i) const { return i / 32 < w.size() && ((w[i / 32] >> (i % 32)) & 1); }
    // bits [lo, lo+cnt) as an integer, cnt <= 32
    uint32_t bits_at(size_t lo, unsigned cnt) const {
        uint64_t v = 0; size_t i = lo / 32; unsigned sh = lo % 32;
        if (i < w.size()) v = w[i];
        if (i + 1 < w.size()) v |= (uint64_t)w[i + 1] << 32;
        v >>= sh;
        return cnt >= 32 ? (uint32_t)v : (uint32_t)(v & ((1ull << cnt) - 1));
    }
    static int cmp(const BigInt& a, const BigInt& b) {
        if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
        for (size_t i = a.w.size(); i-- > 0;) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
        return 0;
    }
    bool operator==(const BigInt& o) const { return cmp(*this, o) == 0; }
    bool operator<(const BigInt& o) const { return cmp(*this, o) < 0; }

    static BigInt add(const BigInt& a, const BigInt& b) {
        BigInt r; size_t n = std::max(a.w.size(), b.w.size()); r.w.resize(n + 1);
        uint64_t c = 0;
        for (size_t i = 0; i < n; i++) {
            c += (uint64_t)(i < a.w.size() ? a.w[i] : 0) + (i < b.w.size() ? b.w[i] : 0);
            r.w[i] = (uint32_t)c; c >>= 32;
        }
        r.w[n] = (uint32_t)c; r.trim(); return r;
    }
    // a - b, requires a >= b
    static BigInt sub(const BigInt& a, const BigInt& b) {
        if (cmp(a, b) < 0) throw std::runtime_error("BigInt::sub underflow");
        BigInt r; r.w.resize(a.w.size());
        int64_t c = 0;
        for (size_t i = 0; i < a.w.size(); i++) {
            c += (int64_t)a.w[i] - (i < b.w.size() ? b.w[i] : 0);
            r.w[i] = (uint32_t)c; c >>= 32;
        }
        r.trim(); return r;
    }
    static BigInt mul(const BigInt& a, const BigInt& b) {
        BigInt r; if (a.is_zero() || b.is_zero()) return r;
        r.w.assign(a.w.size() + b.w.size(), 0);
        for (size_t i = 0; i < a.w.size(); i++) {
            uint64_t c = 0;
            for (size_t j = 0; j < b.w.size(); j++) {
                c += (uint64_t)a.w[i] * b.w[j] + r.w[i + j];
                r.w[i + j] = (uint32_t)c; c >>= 32;
            }
            r.w[i + b.w.size()] = (uint32_t)c;
        }
        r.trim(); return r;
    }
    static BigInt shl(const BigInt& a, size_t s) {
        BigInt r; if (a.is_zero()) return r;
        size_t ws = s / 32; unsigned bs = s % 32;
        r.w.assign(a.w.size() + ws + 1, 0);
        for (size_t i = 0; i < a.w.size(); i++) {
            uint64_t v = (uint64_t)a.w[i] << bs;
            r.w[i + ws] |= (uint32_t)v; r.w[i + ws + 1] |= (uint32_t)(v >> 32);
        }
        r.trim(); return r;
    }
    static BigInt shr(const BigInt& a, size_t s) {
        BigInt r; size_t ws = s / 32; unsigned bs = s % 32;
        if (ws >= a.w.size()) return r;
        r.w.assign(a.w.size() - ws, 0);
        for (size_t i = ws; i < a.w.size(); i++) {
            uint64_t v = a.w[i]; if (i + 1 < a.w.size()) v |= (uint64_t)a.w[i + 1] << 32;
            r.w[i - ws] = (uint32_t)(v >> bs);
        }
        r.trim(); return r;
    }
    static BigInt pow2(size_t e) { return shl(BigInt(1), e); }
    // floor division by shift-subtract on normalised operands (Knuth D would be faster; setup is one-off)
    static void divmod(const BigInt& a, const BigInt& b, BigInt& q, BigInt& r) {
        if (b.is_zero()) throw std::runtime_error("BigInt::divmod by zero");
        q = BigInt(); r = BigInt();
        if (cmp(a, b) < 0) { r = a; return; }
        // Knuth algorithm D, base 2^32
        size_t n = b.w.size(), m = a.w.size() - n;
        if (n == 1) {
            q.w.assign(a.w.size(), 0); uint64_t rem = 0;
            for (size_t i = a.w.size(); i-- > 0;) { uint64_t cur = (rem << 32) | a.w[i]; q.w[i] = (uint32_t)(cur / b.w[0]); rem = cur % b.w[0]; }
            q.trim(); r = BigInt(rem); return;
        }
        unsigned s = 0; { uint32_t t = b.w.back(); while (!(t & 0x80000000u)) { t <<= 1; s++; } }
        BigInt v = shl(b, s), u = shl(a, s);
        u.w.resize(a.w.size() + 1, 0); v.w.resize(n, 0);
        q.w.assign(m + 1, 0);
        for (size_t j = m + 1; j-- > 0;) {
            uint64_t num = ((uint64_t)u.w[j + n] << 32) | u.w[j + n - 1];
            uint64_t qhat = num / v.w[n - 1], rhat = num % v.w[n - 1];
            while (qhat >= (1ull << 32) || qhat * v.w[n - 2] > ((rhat << 32) | u.w[j + n - 2])) {
                qhat--; rhat += v.w[n - 1]; if (rhat >= (1ull << 32)) break;
            }
            int64_t borrow = 0; uint64_t carry = 0;
            for (size_t i = 0; i < n; i++) {
                uint64_t p = qhat * v.w[i] + carry; carry = p >> 32;
                int64_t t = (int64_t)u.w[i + j] - (int64_t)(uint32_t)p + borrow;
                u.w[i + j] = (uint32_t)t; borrow = t >> 32;
            }
            int64_t t = (int64_t)u.w[j + n] - (int64_t)carry + borrow;
            u.w[j + n] = (uint32_t)t; borrow = t >> 32;
            if (borrow < 0) {
                qhat--; uint64_t c = 0;
                for (size_t i = 0; i < n; i++) { c += (uint64_t)u.w[i + j] + v.w[i]; u.w[i + j] = (uint32_t)c; c >>= 32; }
                u.w[j + n] += (uint32_t)c;
            }
            q.w[j] = (uint32_t)qhat;
        }
        q.trim(); u.trim(); r = shr(u, s);
    }
    static BigInt div(const BigInt& a, const BigInt& b) { BigInt q, r; divmod(a, b, q, r); return q; }
    static BigInt mod(const BigInt& a, const BigInt& b) { BigInt q, r; divmod(a, b, q, r); return r; }
};

}  // namespace pb200
