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
