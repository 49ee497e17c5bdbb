// paillier_chip_host.hpp — C++ host mirror of the reference's Rust API for this path, over the C ABI of paillier_b200.h.
//
// The reference is a Rust crate (toolchain absent here), so the host side above the C ABI is C++ with the reference's own names,
// argument meaning and error behaviour (/root/reference/src/paillier.rs):
//     EncryptionPublicKeyAssigned{n, g}                                   :6-9
//     PaillierChip::{construct, get_biguint, encrypt, add}                :11-85
//     paillier_enc_native, paillier_add_native                            :87-97
//     BigUintChip::{construct, assign_integer, assert_equal_fresh}        (biguint-halo2, as called at :134-164, :205-237)
//     base_test().k(..).lookup_bits(..).expect_satisfied(true).run(..)    (halo2-base test harness, :167-181)
// Every arithmetic value — ciphertexts, (q, rem) of every mul_mod, every advice cell — comes from the GPU through the C ABI; this
// file only walks the chip's assignment order and re-checks the constraints the chip would impose
// (`check_constraints`, the MockProver stand-in).  There is no CPU fallback: without a CUDA device every call returns an error.
// Header-only; link with -lpaillier_b200.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "paillier_b200.h"

namespace paillier_halo2 {

// ---- num_bigint::BigUint stand-in (only what the API and the checker need) ---------------------------------------------
struct BigUint {
    std::vector<uint64_t> w;   // little-endian, no leading zero words
    BigUint() {}
    BigUint(uint64_t v) { if (v) w.push_back(v); }
    static BigUint from_words(const uint64_t* p, size_t n) { BigUint r; r.w.assign(p, p + n); r.trim(); return r; }
    void to_words(uint64_t* p, size_t n) const {
        if (w.size() > n) throw std::runtime_error("BigUint::to_words: value too wide");
        for (size_t i = 0; i < n; i++) p[i] = i < w.size() ? w[i] : 0;
    }
    std::vector<uint64_t> words(size_t n) const { std::vector<uint64_t> v(n); to_words(v.data(), n); return v; }
    void trim() { while (!w.empty() && w.back() == 0) w.pop_back(); }
    bool is_zero() const { return w.empty(); }
    size_t bits() const { return w.empty() ? 0 : 64 * (w.size() - 1) + (64 - __builtin_clzll(w.back())); }
    bool bit(size_t i) const { return i / 64 < w.size() && ((w[i / 64] >> (i % 64)) & 1); }
    static int cmp(const BigUint& a, const BigUint& b) {
        if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
        for (size_t i = a.w.size(); i-- > 0;) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
        return 0;
    }
    bool operator==(const BigUint& o) const { return cmp(*this, o) == 0; }
    bool operator!=(const BigUint& o) const { return cmp(*this, o) != 0; }
    bool operator<(const BigUint& o) const { return cmp(*this, o) < 0; }
    friend BigUint operator+(const BigUint& a, const BigUint& b) {
        BigUint r; size_t n = std::max(a.w.size(), b.w.size()); r.w.resize(n + 1);
        unsigned __int128 c = 0;
        for (size_t i = 0; i < n; i++) { c += (unsigned __int128)(i < a.w.size() ? a.w[i] : 0) + (i < b.w.size() ? b.w[i] : 0); r.w[i] = (uint64_t)c; c >>= 64; }
        r.w[n] = (uint64_t)c; r.trim(); return r;
    }
    friend BigUint operator*(const BigUint& a, const BigUint& b) {
        BigUint r; if (a.is_zero() || b.is_zero()) return r;
        r.w.assign(a.w.size() + b.w.size(), 0);
        for (size_t i = 0; i < a.w.size(); i++) {
            unsigned __int128 c = 0;
            for (size_t j = 0; j < b.w.size(); j++) { c += (unsigned __int128)a.w[i] * b.w[j] + r.w[i + j]; r.w[i + j] = (uint64_t)c; c >>= 64; }
            r.w[i + b.w.size()] = (uint64_t)c;
        }
        r.trim(); return r;
    }
    friend BigUint operator<<(const BigUint& a, size_t s) {
        BigUint r; if (a.is_zero()) return r;
        size_t ws = s / 64; unsigned bs = s % 64;
        r.w.assign(a.w.size() + ws + 1, 0);
        for (size_t i = 0; i < a.w.size(); i++) { r.w[i + ws] |= a.w[i] << bs; if (bs) r.w[i + ws + 1] |= a.w[i] >> (64 - bs); }
        r.trim(); return r;
    }
    friend BigUint operator>>(const BigUint& a, size_t s) {
        BigUint r; size_t ws = s / 64; unsigned bs = s % 64;
        if (ws >= a.w.size()) return r;
        r.w.assign(a.w.size() - ws, 0);
        for (size_t i = ws; i < a.w.size(); i++) { r.w[i - ws] = a.w[i] >> bs; if (bs && i + 1 < a.w.size()) r.w[i - ws] |= a.w[i + 1] << (64 - bs); }
        r.trim(); return r;
    }
    BigUint low_bits(size_t k) const {
        BigUint r; r.w.assign(w.begin(), w.begin() + std::min(w.size(), (k + 63) / 64));
        if (k % 64 && r.w.size() == (k + 63) / 64) r.w.back() &= (~0ull) >> (64 - k % 64);
        r.trim(); return r;
    }
    size_t count_ones() const { size_t c = 0; for (uint64_t x : w) c += (size_t)__builtin_popcountll(x); return c; }
};

// ---- halo2 stand-ins ----------------------------------------------------------------------------------------------------------
typedef std::array<uint64_t, 4> Fr;     // BN254 scalar, canonical little-endian words (the values never wrap the field)
inline BigUint fe_to_biguint(const Fr& f) { return BigUint::from_words(f.data(), 4); }

struct Error { int status; std::string what; };
template <class T>
struct Result {
    int status = 0; std::string what; T value;
    bool is_ok() const { return status == 0; }
    T unwrap() const { if (status) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: " + what); return value; }
    static Result ok(T v) { Result r; r.value = std::move(v); return r; }
    static Result err(int s, const char* where) { Result r; r.status = s; r.what = std::string(where) + ": " + pb200_strerror(s); return r; }
};

struct AssignedBigUint {
    std::vector<BigUint> limb_values;   // little-endian limbs as assigned (AssignedBigUint::limbs())
    BigUint int_value;                  // .value()
    unsigned max_limb_bits = 0;         // int_ref().max_limb_bits
    size_t num_limbs() const { return limb_values.size(); }
    const std::vector<BigUint>& limbs() const { return limb_values; }
    const BigUint& value() const { return int_value; }
    AssignedBigUint extend_limbs(size_t k) const { AssignedBigUint r = *this; r.limb_values.resize(limb_values.size() + k); return r; }
};

struct MulModGroup { size_t first_cell; AssignedBigUint a, b, n; };            // what the checker needs to re-derive a group's constraints
struct AssignRecord { size_t first_cell; size_t n_limbs; BigUint value; };
struct N2Record { size_t first_cell; AssignedBigUint n; };

struct Context {
    std::vector<Fr> cells;                      // advice values in assignment order
    std::vector<MulModGroup> mul_mods;
    std::vector<AssignRecord> assigns;
    std::vector<N2Record> n2s;
    std::vector<std::string> failures;          // constraints that do not hold (assert_equal_fresh, checker)
    Fr load_zero() { cells.push_back(Fr{0, 0, 0, 0}); return cells.back(); }
    void append(const uint64_t* words, size_t n_cells) { for (size_t i = 0; i < n_cells; i++) cells.push_back(Fr{words[4 * i], words[4 * i + 1], words[4 * i + 2], words[4 * i + 3]}); }
};

struct RangeChip { uint32_t lookup_bits; };

// one pb200_key per (n, g, enc_bits, limb_bits); keys are cached for the lifetime of the chip
class KeyCache {
public:
    ~KeyCache() { for (auto& kv : keys_) pb200_key_destroy(kv.second); }
    int get(const BigUint& n, const BigUint& g, uint32_t enc_bits, uint32_t limb_bits, pb200_key** out) {
        auto id = std::make_pair(std::make_pair(n.w, g.w), std::make_pair(enc_bits, limb_bits));
        auto it = keys_.find(id);
        if (it != keys_.end()) { *out = it->second; return PB200_OK; }
        const size_t wi = PB200_WORDS(enc_bits);
        if (n.bits() > enc_bits || g.bits() > enc_bits) return PB200_ERR_RANGE;
        std::vector<uint64_t> nw = n.words(wi), gw = g.words(wi);
        int rc = pb200_key_create(0, enc_bits, limb_bits, nw.data(), gw.data(), out);
        if (rc == PB200_OK) keys_[id] = *out;
        return rc;
    }
private:
    std::map<std::pair<std::pair<std::vector<uint64_t>, std::vector<uint64_t>>, std::pair<uint32_t, uint32_t>>, pb200_key*> keys_;
};

class BigUintChip {
public:
    const RangeChip* range; uint32_t limb_bits;
    std::shared_ptr<KeyCache> keys = std::make_shared<KeyCache>();
    static BigUintChip construct(const RangeChip* range, uint32_t limb_bits) { BigUintChip c; c.range = range; c.limb_bits = limb_bits; return c; }

    // assign_integer: limbs and their range-check chunks are cut on the GPU (pb200_assign_cells_batch)
    Result<AssignedBigUint> assign_integer(Context* ctx, const BigUint& v, uint32_t bit_len) const {
        if (bit_len % limb_bits != 0) return Result<AssignedBigUint>::err(PB200_ERR_INVALID_ARG, "assign_integer: bit_len % limb_bits != 0");
        if (v.bits() > bit_len) return Result<AssignedBigUint>::err(PB200_ERR_RANGE, "assign_integer: value does not fit bit_len");
        pb200_key* util = nullptr;
        int rc = utility_key(&util); if (rc) return Result<AssignedBigUint>::err(rc, "assign_integer");
        pb200_cell_layout lay; rc = pb200_cells_layout(util, range->lookup_bits, &lay); if (rc) return Result<AssignedBigUint>::err(rc, "pb200_cells_layout");
        const size_t nl = bit_len / limb_bits, n_cells = nl * lay.cells_per_limb;
        std::vector<uint64_t> vw = v.words(PB200_WORDS(bit_len)), cells(n_cells * 4);
        rc = pb200_assign_cells_batch(util, vw.data(), 1, bit_len, range->lookup_bits, 0, cells.data());
        if (rc) return Result<AssignedBigUint>::err(rc, "pb200_assign_cells_batch");
        AssignedBigUint a; a.int_value = v; a.max_limb_bits = limb_bits;
        for (size_t l = 0; l < nl; l++) a.limb_values.push_back(BigUint::from_words(&cells[4 * l * lay.cells_per_limb], 4));
        ctx->assigns.push_back(AssignRecord{ctx->cells.size(), nl, v});
        ctx->append(cells.data(), n_cells);
        return Result<AssignedBigUint>::ok(a);
    }
    Result<bool> assert_equal_fresh(Context* ctx, const AssignedBigUint& a, const AssignedBigUint& b) const {
        bool eq = a.num_limbs() == b.num_limbs();
        for (size_t i = 0; eq && i < a.num_limbs(); i++) eq = a.limb_values[i] == b.limb_values[i];
        if (!eq) ctx->failures.push_back("assert_equal_fresh");
        return Result<bool>::ok(eq);
    }
    int utility_key(pb200_key** out) const {       // any valid key of this limb width serves the pure formatting entry points
        BigUint n = (BigUint(1) << (2 * limb_bits - 1)) + BigUint(1);
        return keys->get(n, BigUint(2), 2 * limb_bits, limb_bits, out);
    }
};

struct EncryptionPublicKeyAssigned { AssignedBigUint n, g; };

class PaillierChip {
public:
    const BigUintChip* biguint; size_t enc_bits;
    static PaillierChip construct(const BigUintChip* biguint, size_t enc_bits) { PaillierChip c; c.biguint = biguint; c.enc_bits = enc_bits; return c; }

    // src/paillier.rs:22-30 — fold the limbs most significant first
    BigUint get_biguint(const AssignedBigUint& a) const {
        BigUint acc;
        for (size_t i = a.limbs().size(); i-- > 0;) acc = (acc << a.max_limb_bits) + a.limbs()[i];
        return acc;
    }

    // src/paillier.rs:32-60
    Result<AssignedBigUint> encrypt(Context* ctx, const EncryptionPublicKeyAssigned& pk_enc, const AssignedBigUint& m, const AssignedBigUint& r) const {
        typedef Result<AssignedBigUint> R;
        const BigUint n = get_biguint(pk_enc.n), g = get_biguint(pk_enc.g), mv = get_biguint(m), rv = get_biguint(r);
        pb200_key* key = nullptr;
        int rc = biguint->keys->get(n, g, (uint32_t)enc_bits, biguint->limb_bits, &key); if (rc) return R::err(rc, "PaillierChip::encrypt: key");
        AssignedBigUint n2; rc = square_refresh(ctx, key, pk_enc.n, &n2); if (rc) return R::err(rc, "PaillierChip::encrypt: n^2");
        ctx->load_zero();
        const size_t wo = pb200_key_words_out(key), wi = pb200_key_words_in(key);
        // the (q, rem) stream of this unit and the per-key g-chain squarings, from the GPU
        std::vector<std::pair<BigUint, BigUint>> unit, gch;
        {
            std::vector<uint64_t> mw = mv.words(wi), rw = rv.words(wi), c(wo);
            struct Sink { std::vector<std::pair<BigUint, BigUint>>* out; size_t wo; } sink{&unit, wo};
            auto fn = [](void* user, const pb200_witness_chunk* ch) -> int {
                Sink* s = (Sink*)user;
                for (uint64_t i = ch->offsets[0]; i < ch->offsets[ch->n_units]; i++)
                    s->out->push_back({BigUint::from_words(ch->records + i * 2 * s->wo, s->wo), BigUint::from_words(ch->records + i * 2 * s->wo + s->wo, s->wo)});
                return 0;
            };
            rc = pb200_encrypt_witness_batch(key, mw.data(), rw.data(), 1, c.data(), 0, fn, &sink); if (rc) return R::err(rc, "pb200_encrypt_witness_batch");
            std::vector<uint64_t> gw((size_t)enc_bits * 2 * wo);
            rc = pb200_key_g_chain(key, gw.data()); if (rc) return R::err(rc, "pb200_key_g_chain");
            for (size_t i = 0; i < enc_bits; i++) gch.push_back({BigUint::from_words(&gw[i * 2 * wo], wo), BigUint::from_words(&gw[i * 2 * wo + wo], wo)});
        }
        // walk pow_mod_fixed_exp(g, m), pow_mod_fixed_exp(r, n), mul_mod(gm, rn) in the chip's order
        struct G { BigUint a, b, q, rem; };
        std::vector<G> groups; std::vector<size_t> chain_starts;
        size_t ui = 0;
        BigUint gm, rn;
        for (int which = 0; which < 2; which++) {
            const BigUint& base = which ? rv : g; const BigUint& e = which ? n : mv;
            chain_starts.push_back(groups.size());
            BigUint acc(1), sq = base;
            for (size_t i = 0; i < e.bits(); i++) {
                const BigUint cur = sq;
                std::pair<BigUint, BigUint> rec = which ? unit.at(ui++) : gch.at(i);
                groups.push_back(G{cur, cur, rec.first, rec.second}); sq = rec.second;
                if (e.bit(i)) { rec = unit.at(ui++); groups.push_back(G{acc, cur, rec.first, rec.second}); acc = rec.second; }
            }
            (which ? rn : gm) = acc;
        }
        { auto rec = unit.at(ui++); groups.push_back(G{gm, rn, rec.first, rec.second}); }
        if (ui != unit.size()) return R::err(PB200_ERR_INVALID_ARG, "PaillierChip::encrypt: witness stream length");
        rc = assign_groups(ctx, key, n2, groups.size(), [&](size_t i, int f) -> const BigUint& { const G& x = groups[i]; return f == 0 ? x.a : f == 1 ? x.b : f == 2 ? x.q : x.rem; }, chain_starts);
        if (rc) return R::err(rc, "pb200_mulmod_cells_batch");
        return R::ok(fresh(groups.back().rem, n2.num_limbs()));
    }

    // src/paillier.rs:62-85
    Result<AssignedBigUint> add(Context* ctx, const EncryptionPublicKeyAssigned& pk_enc, const AssignedBigUint& c1, const AssignedBigUint& c2) const {
        typedef Result<AssignedBigUint> R;
        const BigUint n = get_biguint(pk_enc.n), g = get_biguint(pk_enc.g), a = get_biguint(c1), b = get_biguint(c2);
        pb200_key* key = nullptr;
        int rc = biguint->keys->get(n, g, (uint32_t)enc_bits, biguint->limb_bits, &key); if (rc) return R::err(rc, "PaillierChip::add: key");
        AssignedBigUint n2; rc = square_refresh(ctx, key, pk_enc.n, &n2); if (rc) return R::err(rc, "PaillierChip::add: n^2");
        ctx->load_zero();
        const size_t wo = pb200_key_words_out(key);
        std::vector<uint64_t> aw = a.words(wo), bw = b.words(wo), out(wo), q(wo);
        rc = pb200_add_batch(key, aw.data(), bw.data(), (uint32_t)wo, 1, out.data(), q.data()); if (rc) return R::err(rc, "pb200_add_batch");
        const BigUint qv = BigUint::from_words(q.data(), wo), rem = BigUint::from_words(out.data(), wo);
        rc = assign_groups(ctx, key, n2, 1, [&](size_t, int f) -> const BigUint& { return f == 0 ? a : f == 1 ? b : f == 2 ? qv : rem; }, {});
        if (rc) return R::err(rc, "pb200_mulmod_cells_batch");
        return R::ok(fresh(rem, n2.num_limbs()));
    }

private:
    AssignedBigUint fresh(const BigUint& v, size_t nl) const {
        AssignedBigUint a; a.int_value = v; a.max_limb_bits = biguint->limb_bits;
        for (size_t l = 0; l < nl; l++) a.limb_values.push_back((v >> (l * biguint->limb_bits)).low_bits(biguint->limb_bits));
        return a;
    }
    // square + refresh of n (src/paillier.rs:39-45): cells from pb200_key_n2_cells, value from pb200_key_n2
    int square_refresh(Context* ctx, pb200_key* key, const AssignedBigUint& n_as, AssignedBigUint* n2) const {
        pb200_cell_layout lay; int rc = pb200_cells_layout(key, biguint->range->lookup_bits, &lay); if (rc) return rc;
        std::vector<uint64_t> cells((size_t)lay.cells_n2 * 4), w(pb200_key_words_out(key));
        rc = pb200_key_n2_cells(key, biguint->range->lookup_bits, 0, cells.data()); if (rc) return rc;
        rc = pb200_key_n2(key, w.data()); if (rc) return rc;
        ctx->n2s.push_back(N2Record{ctx->cells.size(), n_as});
        ctx->append(cells.data(), lay.cells_n2);
        *n2 = fresh(BigUint::from_words(w.data(), w.size()), lay.limbs);
        return PB200_OK;
    }
    int assign_groups(Context* ctx, pb200_key* key, const AssignedBigUint& n2, size_t count, const std::function<const BigUint&(size_t, int)>& field,
                      const std::vector<size_t>& chain_starts) const {
        pb200_cell_layout lay; int rc = pb200_cells_layout(key, biguint->range->lookup_bits, &lay); if (rc) return rc;
        const size_t wo = pb200_key_words_out(key);
        std::vector<uint64_t> in[4];
        for (int f = 0; f < 4; f++) { in[f].resize(count * wo); for (size_t i = 0; i < count; i++) field(i, f).to_words(&in[f][i * wo], wo); }
        std::vector<uint64_t> cells(count * (size_t)lay.cells_per_mulmod * 4);
        rc = pb200_mulmod_cells_batch(key, in[0].data(), in[1].data(), in[2].data(), in[3].data(), count, biguint->range->lookup_bits, 0, cells.data());
        if (rc) return rc;
        for (size_t i = 0; i < count; i++) {
            for (size_t s : chain_starts) if (s == i) ctx->cells.push_back(Fr{1, 0, 0, 0});     // acc = assign_constant(1) of pow_mod_fixed_exp
            ctx->mul_mods.push_back(MulModGroup{ctx->cells.size(), fresh(field(i, 0), lay.limbs), fresh(field(i, 1), lay.limbs), n2});
            ctx->append(&cells[i * (size_t)lay.cells_per_mulmod * 4], lay.cells_per_mulmod);
        }
        return PB200_OK;
    }
};

// src/paillier.rs:87-97 — one-unit calls of the batched GPU entry points; n must be odd (a Paillier modulus) and non-zero
inline Result<BigUint> paillier_enc_native(const BigUint& n, const BigUint& g, const BigUint& m, const BigUint& r, uint32_t enc_bits, KeyCache* keys, uint32_t limb_bits = 64) {
    pb200_key* key = nullptr;
    int rc = keys->get(n, g, enc_bits, limb_bits, &key); if (rc) return Result<BigUint>::err(rc, "paillier_enc_native");
    const size_t wi = pb200_key_words_in(key), wo = pb200_key_words_out(key);
    if (m.bits() > enc_bits || r.bits() > enc_bits) return Result<BigUint>::err(PB200_ERR_RANGE, "paillier_enc_native");
    std::vector<uint64_t> mw = m.words(wi), rw = r.words(wi), c(wo);
    rc = pb200_encrypt_batch(key, mw.data(), rw.data(), 1, c.data()); if (rc) return Result<BigUint>::err(rc, "pb200_encrypt_batch");
    return Result<BigUint>::ok(BigUint::from_words(c.data(), wo));
}
inline Result<BigUint> paillier_add_native(const BigUint& n, const BigUint& c1, const BigUint& c2, uint32_t enc_bits, KeyCache* keys, uint32_t limb_bits = 64) {
    pb200_key* key = nullptr;
    int rc = keys->get(n, BigUint(2), enc_bits, limb_bits, &key); if (rc) return Result<BigUint>::err(rc, "paillier_add_native");
    const size_t wo = pb200_key_words_out(key);
    if (c1.bits() > 64 * wo || c2.bits() > 64 * wo) return Result<BigUint>::err(PB200_ERR_RANGE, "paillier_add_native");
    std::vector<uint64_t> a = c1.words(wo), b = c2.words(wo), out(wo);
    rc = pb200_add_batch(key, a.data(), b.data(), (uint32_t)wo, 1, out.data(), nullptr); if (rc) return Result<BigUint>::err(rc, "pb200_add_batch");
    return Result<BigUint>::ok(BigUint::from_words(out.data(), wo));
}

inline BigUint limb_max(uint32_t limb_bits) {     // B - 1
    BigUint f;
    f.w.assign((limb_bits + 63) / 64, ~0ull);
    if (limb_bits % 64) f.w.back() = (~0ull) >> (64 - limb_bits % 64);
    return f;
}

// ---- the MockProver stand-in: re-derive every constraint BigUintChip imposes on the recorded cells ------------------------------
// (SURVEY.md A.1-A.4, A.6).  Returns the list of violated constraints (empty = satisfied).
inline std::vector<std::string> check_constraints(const Context& ctx, uint32_t limb_bits, uint32_t lookup_bits) {
    std::vector<std::string> bad = ctx.failures;
    auto cell = [&](size_t i) { return fe_to_biguint(ctx.cells.at(i)); };
    const BigUint B1 = (BigUint(1) << limb_bits);
    auto chunks_of = [&](uint32_t bits) { return lookup_bits ? (bits + lookup_bits - 1) / lookup_bits : 0u; };
    auto extra_of = [&](uint32_t bits) { return lookup_bits && bits % lookup_bits ? 1u : 0u; };
    // range_check(v, bits): chunks recompose v, every chunk < 2^lookup_bits, the shifted top chunk as well
    auto range_ok = [&](const BigUint& v, size_t first_chunk, uint32_t bits) {
        if (v.bits() > bits) return false;
        if (!lookup_bits) return true;
        const uint32_t k = chunks_of(bits);
        BigUint acc;
        for (uint32_t j = 0; j < k; j++) { BigUint c = cell(first_chunk + j); if (c.bits() > lookup_bits) return false; acc = acc + (c << (j * lookup_bits)); }
        if (acc != v) return false;
        if (extra_of(bits)) { BigUint x = cell(first_chunk + k); if (x != (cell(first_chunk + k - 1) << (lookup_bits - bits % lookup_bits)) || x.bits() > lookup_bits) return false; }
        return true;
    };
    const uint32_t cpl = 1 + chunks_of(limb_bits) + extra_of(limb_bits);
    auto compose = [&](const std::vector<BigUint>& limbs) { BigUint v; for (size_t i = limbs.size(); i-- > 0;) v = (v << limb_bits) + limbs[i]; return v; };
    for (const AssignRecord& a : ctx.assigns) {
        std::vector<BigUint> limbs;
        for (size_t l = 0; l < a.n_limbs; l++) { limbs.push_back(cell(a.first_cell + l * cpl)); if (!range_ok(limbs.back(), a.first_cell + l * cpl + 1, limb_bits)) bad.push_back("assign_integer: range_check"); }
        if (compose(limbs) != a.value) bad.push_back("assign_integer: limbs do not compose the value");
    }
    auto columns = [&](const std::vector<BigUint>& x, const std::vector<BigUint>& y) {
        std::vector<BigUint> c(x.size() + y.size() - 1);
        for (size_t i = 0; i < x.size(); i++) for (size_t j = 0; j < y.size(); j++) c[i + j] = c[i + j] + x[i] * y[j];
        return c;
    };
    for (const N2Record& r : ctx.n2s) {                               // square + refresh (A.2, A.3)
        const std::vector<BigUint>& nl = r.n.limbs();
        std::vector<BigUint> cols = columns(nl, nl), x = cols;
        size_t c = r.first_cell;
        for (size_t i = 0; i < cols.size(); i++) if (cell(c++) != cols[i]) bad.push_back("square: column");
        // RefreshAux::new(limb_bits, k, k): worst-case spill of every column (all limbs B - 1)
        std::vector<BigUint> full(nl.size(), limb_max(limb_bits));
        std::vector<BigUint> wv = columns(full, full);
        std::vector<unsigned> inc;
        for (size_t i = 0; i < wv.size(); i++) {
            BigUint carry = wv[i] >> limb_bits; unsigned cnt = 0; size_t k = 1;
            while (!carry.is_zero()) { if (i + k >= wv.size()) wv.push_back(BigUint()); wv[i + k] = wv[i + k] + carry.low_bits(limb_bits); carry = carry >> limb_bits; cnt++; k++; }
            wv[i] = wv[i].low_bits(limb_bits); inc.push_back(cnt);
        }
        x.resize(inc.size());
        for (size_t i = 0; i < cols.size(); i++) {
            BigUint limb = x[i];
            for (unsigned j = 0; j <= inc[i]; j++) {
                const BigUint q = cell(c++), rr = cell(c++);
                if (q * B1 + rr != limb || !(rr < B1)) bad.push_back("refresh: div_mod_unsafe");
                if (j == 0) x[i] = rr; else x[i + j] = x[i + j] + rr;
                limb = q;
            }
            if (!limb.is_zero()) bad.push_back("refresh: final carry");
        }
        for (size_t i = 0; i < x.size(); i++) { if (!range_ok(x[i], c, limb_bits)) bad.push_back("refresh: range_check"); c += chunks_of(limb_bits) + extra_of(limb_bits); }
        if (compose(x) != compose(nl) * compose(nl)) bad.push_back("refresh: value");
    }
    for (const MulModGroup& g : ctx.mul_mods) {                      // mul_mod (A.4) + is_equal_muled (A.6)
        const size_t L = g.n.num_limbs(), NC = 2 * L - 1;
        size_t c = g.first_cell;
        std::vector<BigUint> q, rem;
        for (int which = 0; which < 2; which++)
            for (size_t l = 0; l < L; l++) { BigUint v = cell(c); if (!range_ok(v, c + 1, limb_bits)) bad.push_back("mul_mod: q/rem range_check"); (which ? rem : q).push_back(v); c += cpl; }
        std::vector<BigUint> ab = columns(g.a.limbs(), g.b.limbs()), qn = columns(q, g.n.limbs()), qp(NC);
        for (size_t i = 0; i < NC; i++) if (cell(c++) != ab[i]) bad.push_back("mul_mod: ab column");
        for (size_t i = 0; i < NC; i++) if (cell(c++) != qn[i]) bad.push_back("mul_mod: qn column");
        for (size_t i = 0; i < NC; i++) { qp[i] = i < L ? qn[i] + rem[i] : qn[i]; if (cell(c++) != qp[i]) bad.push_back("mul_mod: qn + rem"); }
        const BigUint bmax = limb_max(limb_bits);
        const BigUint word_max = BigUint((uint64_t)L) * bmax * bmax + bmax;
        const uint32_t carry_bits = (uint32_t)(word_max << 1).bits() - limb_bits;
        BigUint carry, acc_extra;
        bool eq = true;
        for (size_t i = 0; i < NC; i++) {
            const BigUint carry_next = cell(c), cs = cell(c + 1), q_acc = cell(c + 2), mod_acc = cell(c + 3);
            if (ab[i] + carry + word_max != qp[i] + carry_next * B1 + cs || !(cs < B1)) bad.push_back("is_equal_muled: carry equation");
            acc_extra = acc_extra + word_max;
            if (q_acc * B1 + mod_acc != acc_extra || !(mod_acc < B1)) bad.push_back("is_equal_muled: acc_extra");
            eq = eq && cs == mod_acc;
            acc_extra = q_acc;
            c += 4;
            if (i + 1 < NC) { if (!range_ok(carry_next, c, carry_bits)) bad.push_back("is_equal_muled: carry range_check"); c += chunks_of(carry_bits) + extra_of(carry_bits); }
            else eq = eq && carry_next == acc_extra;
            carry = carry_next;
        }
        if (cell(c) != BigUint(eq ? 1 : 0) || !eq) bad.push_back("mul_mod: ab != q*n + rem");
    }
    return bad;
}

// ---- base_test() harness (halo2-base utils::testing), enough of it for the reference's tests -------------------------------------
struct BaseTest {
    uint32_t k_ = 0, lookup_bits_ = 0; bool expect_ = true;
    BaseTest& k(uint32_t v) { k_ = v; return *this; }
    BaseTest& lookup_bits(uint32_t v) { lookup_bits_ = v; return *this; }
    BaseTest& expect_satisfied(bool v) { expect_ = v; return *this; }
    // runs the closure, then the constraint re-checker; returns true when the outcome matches expect_satisfied
    bool run(uint32_t limb_bits, const std::function<void(Context*, const RangeChip*)>& f, std::vector<std::string>* why = nullptr, Context* keep = nullptr) {
        Context ctx; RangeChip range{lookup_bits_};
        f(&ctx, &range);
        std::vector<std::string> bad = check_constraints(ctx, limb_bits, lookup_bits_);
        if (ctx.cells.size() > ((size_t)1 << k_) * 64) bad.push_back("circuit does not fit 2^k rows");
        if (why) *why = bad;
        if (keep) *keep = ctx;
        return bad.empty() == expect_;
    }
};
inline BaseTest base_test() { return BaseTest(); }

}  // namespace paillier_halo2
