// test_paillier_chip.cpp — the reference's own tests (/root/reference/src/paillier.rs:107-260), restated in C++ over
// include/paillier_chip_host.hpp: same flows, same names, same assertions.  Every value the chip assigns comes from the GPU;
// the expected ciphertexts come from an independent CPU implementation (OpenSSL BIGNUM), used here as the checker only.
//   test_paillier_encryption  : ENC_BIT_LEN 128, LIMB_BIT_LEN 64, base_test().k(16).lookup_bits(15).expect_satisfied(true)
//   test_encryption_addition  : ENC_BIT_LEN 264, LIMB_BIT_LEN 88
// plus what the reference cannot express: a tampered witness must be rejected, an even or zero modulus is an error, not a panic.
// Build: g++ -std=c++17 -O1 -Iinclude tests/host/test_paillier_chip.cpp -Lpaillier_halo2_b200 -lpaillier_b200 -lcrypto
#include <cstdio>
#include <cstdlib>
#include <random>
#include <openssl/bn.h>
#include "paillier_chip_host.hpp"

using namespace paillier_halo2;

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); exit(1); } } while (0)

static std::mt19937_64 rng(20261018);
static BigUint gen_biguint(size_t bits) {          // rng.gen_biguint(bits): uniform below 2^bits
    BigUint r; r.w.resize((bits + 63) / 64);
    for (auto& x : r.w) x = rng();
    if (bits % 64) r.w.back() &= (~0ull) >> (64 - bits % 64);
    r.trim(); return r;
}

// independent expected values (checker): OpenSSL
static BIGNUM* to_bn(const BigUint& v) {
    std::vector<unsigned char> be(v.w.size() * 8 + 1, 0);
    for (size_t i = 0; i < v.w.size(); i++) for (int b = 0; b < 8; b++) be[be.size() - 1 - (i * 8 + b)] = (unsigned char)(v.w[i] >> (8 * b));
    return BN_bin2bn(be.data(), (int)be.size(), nullptr);
}
static BigUint from_bn(const BIGNUM* b) {
    std::vector<unsigned char> be(BN_num_bytes(b) + 8, 0);
    int n = BN_bn2bin(b, be.data());
    BigUint r; r.w.assign((n + 7) / 8, 0);
    for (int i = 0; i < n; i++) r.w[(n - 1 - i) / 8] |= (uint64_t)be[i] << (8 * ((n - 1 - i) % 8));
    r.trim(); return r;
}
static BigUint expected_enc(const BigUint& n, const BigUint& g, const BigUint& m, const BigUint& r) {
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM *N = to_bn(n), *G = to_bn(g), *M = to_bn(m), *R = to_bn(r), *N2 = BN_new(), *A = BN_new(), *B = BN_new();
    BN_mul(N2, N, N, ctx); BN_mod_exp(A, G, M, N2, ctx); BN_mod_exp(B, R, N, N2, ctx); BN_mod_mul(A, A, B, N2, ctx);
    BigUint out = from_bn(A);
    BN_free(N); BN_free(G); BN_free(M); BN_free(R); BN_free(N2); BN_free(A); BN_free(B); BN_CTX_free(ctx);
    return out;
}
static BigUint expected_add(const BigUint& n, const BigUint& c1, const BigUint& c2) {
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM *N = to_bn(n), *A = to_bn(c1), *B = to_bn(c2), *N2 = BN_new();
    BN_mul(N2, N, N, ctx); BN_mod_mul(A, A, B, N2, ctx);
    BigUint out = from_bn(A);
    BN_free(N); BN_free(A); BN_free(B); BN_free(N2); BN_CTX_free(ctx);
    return out;
}

// src/paillier.rs:121-165
static void paillier_enc_circuit(Context* ctx, const RangeChip* range, size_t enc_bit_len, size_t limb_bit_len,
                                 BigUint n, BigUint g, BigUint m, BigUint r, BigUint res) {
    BigUintChip biguint_chip = BigUintChip::construct(range, (uint32_t)limb_bit_len);
    PaillierChip paillier_chip = PaillierChip::construct(&biguint_chip, enc_bit_len);
    AssignedBigUint n_assigned = biguint_chip.assign_integer(ctx, n, (uint32_t)enc_bit_len).unwrap();
    AssignedBigUint g_assigned = biguint_chip.assign_integer(ctx, g, (uint32_t)enc_bit_len).unwrap();
    EncryptionPublicKeyAssigned pk_enc{n_assigned, g_assigned};
    AssignedBigUint m_assigned = biguint_chip.assign_integer(ctx, m, (uint32_t)enc_bit_len).unwrap();
    AssignedBigUint r_assigned = biguint_chip.assign_integer(ctx, r, (uint32_t)enc_bit_len).unwrap();
    AssignedBigUint c_assigned = paillier_chip.encrypt(ctx, pk_enc, m_assigned, r_assigned).unwrap();
    AssignedBigUint res_assigned = biguint_chip.assign_integer(ctx, res, (uint32_t)enc_bit_len * 2).unwrap();
    CHECK(c_assigned.value() == res_assigned.value());                                   // assert_eq!(a, b)
    CHECK(biguint_chip.assert_equal_fresh(ctx, c_assigned, res_assigned).unwrap());
}

// src/paillier.rs:191-238
static void paillier_enc_add(Context* ctx, const RangeChip* range, size_t enc_bit_len, size_t limb_bit_len,
                             BigUint n, BigUint g, BigUint c1, BigUint c2, BigUint res) {
    BigUintChip biguint_chip = BigUintChip::construct(range, (uint32_t)limb_bit_len);
    PaillierChip paillier_chip = PaillierChip::construct(&biguint_chip, enc_bit_len);
    AssignedBigUint n_assigned = biguint_chip.assign_integer(ctx, n, (uint32_t)enc_bit_len).unwrap();
    AssignedBigUint g_assigned = biguint_chip.assign_integer(ctx, g, (uint32_t)enc_bit_len).unwrap();
    EncryptionPublicKeyAssigned pk_enc{n_assigned, g_assigned};
    AssignedBigUint c1_assigned = biguint_chip.assign_integer(ctx, c1, (uint32_t)enc_bit_len).unwrap();
    AssignedBigUint c2_assigned = biguint_chip.assign_integer(ctx, c2, (uint32_t)enc_bit_len).unwrap();
    AssignedBigUint c_add_assigned = paillier_chip.add(ctx, pk_enc, c1_assigned, c2_assigned).unwrap();
    AssignedBigUint res_assigned = biguint_chip.assign_integer(ctx, res, (uint32_t)enc_bit_len * 2).unwrap();
    CHECK(c_add_assigned.value() == res_assigned.value());
    CHECK(biguint_chip.assert_equal_fresh(ctx, c_add_assigned, res_assigned).unwrap());
}

static BigUint nonzero(BigUint v) { if (v.is_zero()) v = BigUint(1); return v; }   // num-bigint panics on n = 0; any other n (even, short) is the reference's own test distribution

static void test_paillier_encryption(size_t ENC_BIT_LEN, size_t LIMB_BIT_LEN, int rounds, uint32_t K = 16, uint32_t LOOKUP = 15) {
    for (int it = 0; it < rounds; it++) {
        std::vector<std::string> why; Context kept;
        BigUint n, g, m, r;
        bool ok = base_test().k(K).lookup_bits(LOOKUP).expect_satisfied(true).run((uint32_t)LIMB_BIT_LEN, [&](Context* ctx, const RangeChip* range) {
            n = nonzero(gen_biguint(ENC_BIT_LEN)); g = gen_biguint(ENC_BIT_LEN); m = gen_biguint(ENC_BIT_LEN); r = gen_biguint(ENC_BIT_LEN);
            if (it == 1) m = BigUint();                                  // empty g-chain
            if (it == 2) { r = BigUint(1); m = BigUint(1); }
            BigUint res = expected_enc(n, g, m, r);
            KeyCache keys;
            CHECK(paillier_enc_native(n, g, m, r, (uint32_t)ENC_BIT_LEN, &keys, (uint32_t)LIMB_BIT_LEN).unwrap() == res);
            paillier_enc_circuit(ctx, range, ENC_BIT_LEN, LIMB_BIT_LEN, n, g, m, r, res);
        }, &why, &kept);
        for (auto& w : why) fprintf(stderr, "unsatisfied: %s\n", w.c_str());
        CHECK(ok);
        if (it == 0) {                       // a tampered witness must be rejected by the constraints
            Context bad = kept;
            CHECK(!bad.mul_mods.empty());
            size_t victim = bad.mul_mods[bad.mul_mods.size() / 2].first_cell;
            bad.cells[victim][0] ^= 1;       // lowest limb of q of one mul_mod
            CHECK(!check_constraints(bad, (uint32_t)LIMB_BIT_LEN, LOOKUP).empty());
            Context bad2 = kept;
            bad2.cells[bad2.mul_mods.back().first_cell + 3][0] += 1;     // a range-check chunk
            CHECK(!check_constraints(bad2, (uint32_t)LIMB_BIT_LEN, LOOKUP).empty());
        }
        printf("test_paillier_encryption[%zu/%zu] round %d ok: %zu cells, %zu mul_mod groups\n", ENC_BIT_LEN, LIMB_BIT_LEN, it, kept.cells.size(), kept.mul_mods.size());
    }
}

static void test_encryption_addition(size_t ENC_BIT_LEN, size_t LIMB_BIT_LEN, int rounds) {
    for (int it = 0; it < rounds; it++) {
        std::vector<std::string> why; Context kept;
        bool ok = base_test().k(16).lookup_bits(15).expect_satisfied(true).run((uint32_t)LIMB_BIT_LEN, [&](Context* ctx, const RangeChip* range) {
            BigUint n = nonzero(gen_biguint(ENC_BIT_LEN)), g = gen_biguint(ENC_BIT_LEN), c1 = gen_biguint(ENC_BIT_LEN), c2 = gen_biguint(ENC_BIT_LEN);
            BigUint res = expected_add(n, c1, c2);
            KeyCache keys;
            CHECK(paillier_add_native(n, c1, c2, (uint32_t)ENC_BIT_LEN, &keys, (uint32_t)LIMB_BIT_LEN).unwrap() == res);
            paillier_enc_add(ctx, range, ENC_BIT_LEN, LIMB_BIT_LEN, n, g, c1, c2, res);
        }, &why, &kept);
        for (auto& w : why) fprintf(stderr, "unsatisfied: %s\n", w.c_str());
        CHECK(ok);
        printf("test_encryption_addition[%zu/%zu] round %d ok: %zu cells\n", ENC_BIT_LEN, LIMB_BIT_LEN, it, kept.cells.size());
    }
}

static void test_error_behaviour() {
    KeyCache keys;
    // num-bigint panics on a zero modulus (src/paillier.rs:89-91); here it is an Err, never an abort
    CHECK(paillier_enc_native(BigUint(), BigUint(3), BigUint(5), BigUint(7), 128, &keys).status == PB200_ERR_ZERO_MODULUS);
    CHECK(paillier_enc_native(BigUint(10), BigUint(3), BigUint(5), BigUint(7), 128, &keys).unwrap() == expected_enc(BigUint(10), BigUint(3), BigUint(5), BigUint(7)));   // even n: accepted like the reference
    RangeChip range{15};
    BigUintChip chip = BigUintChip::construct(&range, 64);
    Context ctx;
    CHECK(chip.assign_integer(&ctx, BigUint(5), 100).status == PB200_ERR_INVALID_ARG);          // bit_len % limb_bits
    CHECK(chip.assign_integer(&ctx, BigUint(1) << 130, 128).status == PB200_ERR_RANGE);
    bool threw = false;
    try { chip.assign_integer(&ctx, BigUint(5), 100).unwrap(); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    printf("test_error_behaviour ok\n");
}

int main(int argc, char** argv) {
    if (pb200_device_count() == 0) {
        // no CPU fallback: the host layer reports the failure instead of computing anything
        KeyCache keys;
        auto r = paillier_enc_native(BigUint(11), BigUint(3), BigUint(5), BigUint(7), 128, &keys);
        CHECK(r.status == PB200_ERR_CUDA);
        printf("no CUDA device: paillier_enc_native -> %s\n", r.what.c_str());
        return argc > 1 ? 0 : 2;
    }
    test_error_behaviour();
    test_paillier_encryption(128, 64, 3);
    test_encryption_addition(264, 88, 2);
    test_paillier_encryption(264, 88, 1);       // the reference's second limb shape on the encrypt flow
    test_encryption_addition(128, 64, 1);
    test_paillier_encryption(128, 64, 1, 14, 13);   // the shape of src/bench.rs:161-164 (k = 14, lookup_bits = 13)
    printf("all host tests passed\n");
    return 0;
}
