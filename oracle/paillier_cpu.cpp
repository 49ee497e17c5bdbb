// paillier_cpu.cpp — CPU restatement of the reference's native path on OpenSSL BIGNUM.
// TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ (cross-check of the Python oracle) and by
// bench.py's cpu_baseline / `--impl reference` legs.  Never linked into the product library.
//
// Restates /root/reference/src/paillier.rs:87-92 (paillier_enc_native: n2 = n*n; g.modpow(m,n2);
// r.modpow(n,n2); (gm*rn) % n2) and :94-97 (paillier_add_native: (c1*c2) % n2).  The reference
// uses num-bigint 0.4.4 (Cargo.toml:12), which cannot be built here (no Rust toolchain, SURVEY.md
// §0); OpenSSL's BN_mod_exp (Montgomery, sliding window, assembly kernels) is the same exact
// integer function and is faster than num-bigint, i.e. generous to the reference.  kind = "port".
//
// Words are little-endian uint64 (same layout as include/paillier_b200.h).
#include <openssl/bn.h>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include <dlfcn.h>

static BIGNUM* from_words(const uint64_t* w, int n) {
    std::vector<unsigned char> buf((size_t)n * 8);
    memcpy(buf.data(), w, buf.size());
    return BN_lebin2bn(buf.data(), (int)buf.size(), nullptr);
}
static void to_words(const BIGNUM* b, uint64_t* w, int n) {
    std::vector<unsigned char> buf((size_t)n * 8);
    BN_bn2lebinpad(b, buf.data(), (int)buf.size());
    memcpy(w, buf.data(), buf.size());
}

extern "C" {

int cpu_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// c[u] = g^m[u] * r[u]^n mod n^2 for u in [0, count); `threads` worker threads.  Returns 0 on success.
int cpu_paillier_enc_batch(const uint64_t* n_w, const uint64_t* g_w, int words_in, const uint64_t* m_w,
                           const uint64_t* r_w, size_t count, uint64_t* c_w, int threads) {
    if (threads < 1) threads = 1;
    int words_out = 2 * words_in;
    std::atomic<size_t> next{0};
    std::atomic<int> err{0};
    auto work = [&]() {
        BN_CTX* ctx = BN_CTX_new();
        BIGNUM* n = from_words(n_w, words_in);
        BIGNUM* g = from_words(g_w, words_in);
        BIGNUM* n2 = BN_new(); BIGNUM* gm = BN_new(); BIGNUM* rn = BN_new(); BIGNUM* c = BN_new();
        BN_mul(n2, n, n, ctx);                                   // src/paillier.rs:88
        for (;;) {
            size_t u = next.fetch_add(1);
            if (u >= count) break;
            BIGNUM* m = from_words(m_w + u * words_in, words_in);
            BIGNUM* r = from_words(r_w + u * words_in, words_in);
            if (!BN_mod_exp(gm, g, m, n2, ctx)) err = 1;         // :89
            if (!BN_mod_exp(rn, r, n, n2, ctx)) err = 1;         // :90
            if (!BN_mod_mul(c, gm, rn, n2, ctx)) err = 1;        // :91
            to_words(c, c_w + u * words_out, words_out);
            BN_free(m); BN_free(r);
        }
        BN_free(n); BN_free(g); BN_free(n2); BN_free(gm); BN_free(rn); BN_free(c); BN_CTX_free(ctx);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return err.load();
}

// out[u] = c1[u] * c2[u] mod n^2
int cpu_paillier_add_batch(const uint64_t* n_w, int words_in, const uint64_t* c1_w, const uint64_t* c2_w, int c_words,
                           size_t count, uint64_t* out_w, int threads) {
    if (threads < 1) threads = 1;
    int words_out = 2 * words_in;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        BN_CTX* ctx = BN_CTX_new();
        BIGNUM* n = from_words(n_w, words_in);
        BIGNUM* n2 = BN_new(); BIGNUM* c = BN_new();
        BN_mul(n2, n, n, ctx);                                   // src/paillier.rs:95
        for (;;) {
            size_t u0 = next.fetch_add(256);
            if (u0 >= count) break;
            for (size_t u = u0; u < count && u < u0 + 256; u++) {
                BIGNUM* a = from_words(c1_w + u * c_words, c_words);
                BIGNUM* b = from_words(c2_w + u * c_words, c_words);
                BN_mul(c, a, b, ctx); BN_mod(c, c, n2, ctx);     // :96
                to_words(c, out_w + u * words_out, words_out);
                BN_free(a); BN_free(b);
            }
        }
        BN_free(n); BN_free(n2); BN_free(c); BN_CTX_free(ctx);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return 0;
}

// out = prod c[u] mod n^2 (fold of paillier_add_native); threads fold strided partials, then combined
int cpu_paillier_tally(const uint64_t* n_w, int words_in, const uint64_t* c_w, size_t count, uint64_t* out_w, int threads) {
    if (threads < 1) threads = 1;
    int words_out = 2 * words_in;
    std::vector<std::vector<uint64_t>> partial(threads, std::vector<uint64_t>(words_out, 0));
    auto work = [&](int t) {
        BN_CTX* ctx = BN_CTX_new();
        BIGNUM* n = from_words(n_w, words_in);
        BIGNUM* n2 = BN_new(); BIGNUM* acc = BN_new();
        BN_mul(n2, n, n, ctx); BN_one(acc); BN_mod(acc, acc, n2, ctx);
        for (size_t u = t; u < count; u += threads) {
            BIGNUM* a = from_words(c_w + u * words_out, words_out);
            BN_mod_mul(acc, acc, a, n2, ctx);
            BN_free(a);
        }
        to_words(acc, partial[t].data(), words_out);
        BN_free(n); BN_free(n2); BN_free(acc); BN_CTX_free(ctx);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM* n = from_words(n_w, words_in);
    BIGNUM* n2 = BN_new(); BIGNUM* acc = BN_new();
    BN_mul(n2, n, n, ctx); BN_one(acc); BN_mod(acc, acc, n2, ctx);
    for (int t = 0; t < threads; t++) { BIGNUM* a = from_words(partial[t].data(), words_out); BN_mod_mul(acc, acc, a, n2, ctx); BN_free(a); }
    to_words(acc, out_w, words_out);
    BN_free(n); BN_free(n2); BN_free(acc); BN_CTX_free(ctx);
    return 0;
}

}  // extern "C"

// ---- witness chain (PaillierChip::encrypt's mul_mod sequence, /root/reference/src/paillier.rs:51,55,57) ----------
// Restates BigUintChip::pow_mod_fixed_exp / mul_mod witness generation: LSB-first
// square-and-multiply, every step full = a*b; (q, rem) = full.div_rem(n^2).
// Per-unit stream and digest exactly as include/paillier_b200.h defines them: popcount(m) g-chain mul records (the g-chain
// squarings are per key and not part of the unit stream), the r-chain records (sqr_i then, if bit i of n is set, mul_i; the
// wasted last squaring included), the final gm*rn record.  Record hash H = sum_j w_j C^(j+1) over q's then rem's words;
// D = 0xcbf29ce484222325; D = (D ^ H) * 0x100000001b3 per record.
// Two exact-integer backends, selected by `backend`: 0 = OpenSSL BN_mul/BN_sqr + BN_div (23 us per step at |n| = 2048),
// 1 = GMP mpz_mul + mpz_tdiv_qr (9 us per step; libgmp.so.10 is in the image without headers, so the six entry points used
// are declared here and resolved with dlopen — when that fails the call returns -1 and the caller uses backend 0).
static const uint64_t DIG_INIT = 0xcbf29ce484222325ull, DIG_PRIME = 0x100000001b3ull, DIG_C = 0x9E3779B97F4A7C15ull;

struct BnNum {                                   // OpenSSL backend
    BIGNUM* v; static thread_local BN_CTX* ctx;
    BnNum() : v(BN_new()) {}
    ~BnNum() { BN_free(v); }
    void set_words(const uint64_t* w, int n) { BN_lebin2bn((const unsigned char*)w, n * 8, v); }
    void get_words(uint64_t* w, int n) const { BN_bn2lebinpad(v, (unsigned char*)w, n * 8); }
    void set(const BnNum& o) { BN_copy(v, o.v); }
    void set_one() { BN_one(v); }
    int bits() const { return BN_num_bits(v); }
    bool bit(int i) const { return BN_is_bit_set(v, i); }
    static void mul(BnNum& out, const BnNum& a, const BnNum& b) { if (&a == &b) BN_sqr(out.v, a.v, ctx); else BN_mul(out.v, a.v, b.v, ctx); }
    static void divqr(BnNum& q, BnNum& r, const BnNum& x, const BnNum& d) { BN_div(q.v, r.v, x.v, d.v, ctx); }
    static bool begin_thread() { ctx = BN_CTX_new(); return true; }
    static void end_thread() { BN_CTX_free(ctx); ctx = nullptr; }
};
thread_local BN_CTX* BnNum::ctx = nullptr;

struct GmpApi {                                  // the mpz_t ABI of GMP 6: {int alloc; int size; limb* d}
    struct mpz { int alloc, size; uint64_t* d; };
    void (*init)(mpz*); void (*clear)(mpz*); void (*mul)(mpz*, const mpz*, const mpz*); void (*tdiv_qr)(mpz*, mpz*, const mpz*, const mpz*);
    void (*import_)(mpz*, size_t, int, size_t, int, size_t, const void*); void* (*export_)(void*, size_t*, int, size_t, int, size_t, const mpz*);
    void (*set)(mpz*, const mpz*); void (*set_ui)(mpz*, unsigned long); size_t (*sizeinbase)(const mpz*, int); int (*tstbit)(const mpz*, unsigned long);
    bool ok = false;
    GmpApi() {
        void* h = dlopen("libgmp.so.10", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        auto sym = [&](const char* n) { return dlsym(h, n); };
        init = (decltype(init))sym("__gmpz_init"); clear = (decltype(clear))sym("__gmpz_clear"); mul = (decltype(mul))sym("__gmpz_mul");
        tdiv_qr = (decltype(tdiv_qr))sym("__gmpz_tdiv_qr"); import_ = (decltype(import_))sym("__gmpz_import");
        export_ = (decltype(export_))sym("__gmpz_export"); set = (decltype(set))sym("__gmpz_set"); set_ui = (decltype(set_ui))sym("__gmpz_set_ui");
        sizeinbase = (decltype(sizeinbase))sym("__gmpz_sizeinbase"); tstbit = (decltype(tstbit))sym("__gmpz_tstbit");
        ok = init && clear && mul && tdiv_qr && import_ && export_ && set && set_ui && sizeinbase && tstbit;
    }
};
static GmpApi& gmp() { static GmpApi g; return g; }

struct GmpNum {
    GmpApi::mpz v;
    GmpNum() { gmp().init(&v); }
    ~GmpNum() { gmp().clear(&v); }
    void set_words(const uint64_t* w, int n) { gmp().import_(&v, (size_t)n, -1, 8, 0, 0, w); }
    void get_words(uint64_t* w, int n) const { memset(w, 0, (size_t)n * 8); size_t cnt = 0; if (v.size) gmp().export_(w, &cnt, -1, 8, 0, 0, &v); }
    void set(const GmpNum& o) { gmp().set(&v, &o.v); }
    void set_one() { gmp().set_ui(&v, 1); }
    int bits() const { return v.size ? (int)gmp().sizeinbase(&v, 2) : 0; }
    bool bit(int i) const { return gmp().tstbit(&v, (unsigned long)i) != 0; }
    static void mul(GmpNum& out, const GmpNum& a, const GmpNum& b) { gmp().mul(&out.v, &a.v, &b.v); }
    static void divqr(GmpNum& q, GmpNum& r, const GmpNum& x, const GmpNum& d) { gmp().tdiv_qr(&q.v, &r.v, &x.v, &d.v); }
    static bool begin_thread() { return true; }
    static void end_thread() {}
};

template <class Num>
static int witness_digest_impl(const uint64_t* n_w, const uint64_t* g_w, int words_in, const uint64_t* m_w, const uint64_t* r_w,
                               size_t count, uint64_t* c_w, uint64_t* digest_out, int threads) {
    if (threads < 1) threads = 1;
    const int k = 2 * words_in;
    std::vector<uint64_t> cpow(2 * (size_t)k);
    { uint64_t c = DIG_C; for (size_t j = 0; j < cpow.size(); j++) { cpow[j] = c; c *= DIG_C; } }
    Num::begin_thread();
    Num n0, n2;
    n0.set_words(n_w, words_in);
    Num::mul(n2, n0, n0);                                         // src/paillier.rs:39-45: n^2
    const int ebits = n0.bits();
    // per-key g-chain: gpow[i] = g^(2^i) mod n^2 (entry 0 is g as assigned, not reduced)
    std::vector<Num> gpow((size_t)64 * words_in);
    {
        Num cur, full, q;
        cur.set_words(g_w, words_in);
        for (size_t i = 0; i < gpow.size(); i++) {
            gpow[i].set(cur);
            Num::mul(full, cur, cur);
            Num::divqr(q, cur, full, n2);
        }
    }
    Num::end_thread();
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        Num::begin_thread();
        {
            Num full, q, acc, cur, sq, gm, rem;
            std::vector<uint64_t> tq(k + 1), tr(k);
            for (;;) {
                const size_t u = next.fetch_add(1);
                if (u >= count) break;
                uint64_t D = DIG_INIT;
                auto emit = [&](const Num& qq, const Num& rr) {
                    if (qq.bits() > 64 * k) { bad = 1; return; }
                    qq.get_words(tq.data(), k); rr.get_words(tr.data(), k);
                    uint64_t h = 0;
                    for (int i = 0; i < k; i++) h += tq[i] * cpow[i];
                    for (int i = 0; i < k; i++) h += tr[i] * cpow[k + i];
                    D = (D ^ h) * DIG_PRIME;
                };
                // g-chain (A.5 over the bits of m): acc = 1; for each set bit i: acc = mul_mod(acc, g^(2^i))   — mul records only
                const uint64_t* mw = m_w + u * words_in;
                acc.set_one();
                for (int i = 0; i < 64 * words_in; i++) {
                    if (!((mw[i >> 6] >> (i & 63)) & 1)) continue;
                    Num::mul(full, acc, gpow[i]);
                    Num::divqr(q, acc, full, n2);
                    emit(q, acc);
                }
                gm.set(acc);
                // r-chain (A.5 over the bits of n): cur = sq; sq = square_mod(cur); if bit: acc = mul_mod(acc, cur)
                sq.set_words(r_w + u * words_in, words_in);
                acc.set_one();
                for (int i = 0; i < ebits; i++) {
                    cur.set(sq);
                    Num::mul(full, cur, cur);
                    Num::divqr(q, sq, full, n2);
                    emit(q, sq);
                    if (n0.bit(i)) {
                        Num::mul(full, acc, cur);
                        Num::divqr(q, acc, full, n2);
                        emit(q, acc);
                    }
                }
                Num::mul(full, gm, acc);                          // src/paillier.rs:57
                Num::divqr(q, rem, full, n2);
                emit(q, rem);
                digest_out[u] = D;
                if (c_w) rem.get_words(c_w + u * k, k);
            }
        }
        Num::end_thread();
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return bad.load();
}

extern "C" {
// digest_out[u], c_w[u] (nullable) for u in [0, count).  Returns 0; 1 if some quotient does not fit words_out words (the chip's
// range check on q would fail; that unit's digest is then meaningless); -1 if the requested backend is unavailable.
int cpu_witness_digest_batch(const uint64_t* n_w, const uint64_t* g_w, int words_in, const uint64_t* m_w, const uint64_t* r_w,
                             size_t count, uint64_t* c_w, uint64_t* digest_out, int threads, int backend) {
    if (backend == 1) {
        if (!gmp().ok) return -1;
        return witness_digest_impl<GmpNum>(n_w, g_w, words_in, m_w, r_w, count, c_w, digest_out, threads);
    }
    return witness_digest_impl<BnNum>(n_w, g_w, words_in, m_w, r_w, count, c_w, digest_out, threads);
}

}  // extern "C"
