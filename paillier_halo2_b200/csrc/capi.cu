// capi.cu — the C ABI declared in include/paillier_b200.h.
//
// Host-side glue only: validation (the reference's range checks / panics mapped to status codes),
// per-key constant setup, staging copies and engine dispatch.  All arithmetic per ciphertext runs in
// CUDA kernels (block28_kernels.cu, simple64_kernels.cu); there is no CPU fallback.
#include "../../include/paillier_b200.h"
#include "engine.hpp"
#include "cells.hpp"
#include <cstring>
#include <map>
#include <mutex>
#include <new>

using namespace pb200;

static thread_local std::string t_cuda_error;

static int cuda_fail(cudaError_t e, const char* where) {
    t_cuda_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return PB200_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

struct DevBuf {  // grow-only device buffer
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct pb200_key {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t n_bits = 0, limb_bits = 0, words_in = 0, words_out = 0;
    BigInt n, g, n2;
    SimpleConsts* d_simple = nullptr;
    u64* d_gchain = nullptr;        // lazy: n_bits records
    int* d_flags = nullptr;
    Block28Key* fast = nullptr;
    std::string fast_why;
    int engine = 0;                 // 0 auto, 1 simple64, 2 block28
    DevBuf in_a, in_b, out_a, out_b, scratch, offs;
    std::string engine_name;
    // K4 cell expansion: per lookup_bits layout + device constants (n^2 limbs, word_max, q_acc, mod_acc), refresh spill vector
    struct CellCtx { CellLayout Y; u64* d_consts = nullptr; int* d_inc = nullptr; u64* d_mtab = nullptr; int n_out = 0; int n2_cells = 0; };
    std::map<uint32_t, CellCtx> cells;
    DevBuf cin_a, cin_b, cin_q, cin_r;
    int sms = 148;
};

static bool use_fast(const pb200_key* k) { return k->fast && k->engine != 1; }

extern "C" {

const char* pb200_strerror(int s) {
    switch (s) {
        case PB200_OK: return "ok";
        case PB200_ERR_INVALID_ARG: return "invalid argument";
        case PB200_ERR_ZERO_MODULUS: return "modulus n is zero (num-bigint would panic)";
        case PB200_ERR_EVEN_MODULUS: return "modulus n is even (GPU path requires odd n)";
        case PB200_ERR_RANGE: return "input does not fit its declared bit width (range check fails)";
        case PB200_ERR_UNSUPPORTED: return "key size not supported by any compiled engine";
        case PB200_ERR_CUDA: return "CUDA failure";
        case PB200_ERR_NOMEM: return "out of memory";
        case PB200_ERR_SINK: return "witness sink aborted";
        default: return "unknown status";
    }
}
const char* pb200_last_cuda_error(void) { return t_cuda_error.c_str(); }
const char* pb200_version(void) { return "paillier_b200 0.1 (sm_100a)"; }
int pb200_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
uint64_t pb200_kernel_launches(void) { return g_kernel_launches.load(); }

static bool fits(const uint64_t* v, uint32_t words, uint32_t bits) {
    for (uint32_t i = 0; i < words; i++) {
        uint32_t lo = i * 64;
        if (lo >= bits) { if (v[i]) return false; }
        else if (bits - lo < 64) { if (v[i] >> (bits - lo)) return false; }
    }
    return true;
}

int pb200_key_create(int device, uint32_t n_bits, uint32_t limb_bits, const uint64_t* n_le, const uint64_t* g_le,
                     pb200_key** out) {
    if (!out) return PB200_ERR_INVALID_ARG;
    *out = nullptr;
    if (!n_le || !g_le || n_bits == 0 || limb_bits == 0 || limb_bits > 128) return PB200_ERR_INVALID_ARG;
    if (n_bits % limb_bits != 0) return PB200_ERR_INVALID_ARG;  // assign_integer asserts bit_len % limb_bits == 0
    uint32_t win = PB200_WORDS(n_bits), wout = PB200_WORDS(2 * n_bits);
    if (wout > PB200_SIMPLE_MAXK) return PB200_ERR_UNSUPPORTED;
    if (!fits(n_le, win, n_bits) || !fits(g_le, win, n_bits)) return PB200_ERR_RANGE;
    BigInt n = BigInt::from_u64_le(n_le, win), g = BigInt::from_u64_le(g_le, win);
    if (n.is_zero()) return PB200_ERR_ZERO_MODULUS;
    if (!n.is_odd()) return PB200_ERR_EVEN_MODULUS;
    int ndev = pb200_device_count();
    if (device < 0 || device >= ndev) { t_cuda_error = "no such CUDA device"; return PB200_ERR_CUDA; }
    CU(cudaSetDevice(device));
    pb200_key* k = new (std::nothrow) pb200_key();
    if (!k) return PB200_ERR_NOMEM;
    k->device = device; k->n_bits = n_bits; k->limb_bits = limb_bits; k->words_in = win; k->words_out = wout;
    k->n = n; k->g = g; k->n2 = BigInt::mul(n, n);
    SimpleConsts* h = new SimpleConsts();
    memset(h, 0, sizeof(*h));
    h->k = (int)wout; h->kin = (int)win; h->n_bits = (int)n_bits; h->exp_bits = (int)n.bits();
    h->s = (int)(64 * wout - k->n2.bits());
    BigInt Nt = BigInt::shl(k->n2, h->s);
    BigInt mu = BigInt::div(BigInt::pow2(128 * (size_t)wout), Nt);
    Nt.to_u64_le(h->Nt, wout); mu.to_u64_le(h->mu, wout + 1);
    n.to_u64_le(h->n, win); g.to_u64_le(h->g, win);
    cudaError_t e = cudaStreamCreateWithFlags(&k->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&k->d_simple, sizeof(SimpleConsts));
    if (e == cudaSuccess) e = cudaMalloc(&k->d_flags, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(k->d_simple, h, sizeof(SimpleConsts), cudaMemcpyHostToDevice, k->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(k->d_flags, 0, sizeof(int), k->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(k->stream);
    delete h;
    if (e != cudaSuccess) { int rc = cuda_fail(e, "pb200_key_create"); pb200_key_destroy(k); return rc; }
    cudaError_t fe = cudaSuccess;
    { cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) k->sms = prop.multiProcessorCount; }
    k->fast = block28_create(n, g, n_bits, device, k->stream, &k->fast_why, &fe);
    if (fe != cudaSuccess) { int rc = cuda_fail(fe, "block28_create"); pb200_key_destroy(k); return rc; }
    k->engine_name = k->fast ? block28_name(k->fast) : "simple64";
    *out = k;
    return PB200_OK;
}

void pb200_key_destroy(pb200_key* k) {
    if (!k) return;
    cudaSetDevice(k->device);
    if (k->stream) cudaStreamSynchronize(k->stream);
    if (k->fast) block28_destroy(k->fast);
    if (k->d_simple) cudaFree(k->d_simple);
    if (k->d_gchain) cudaFree(k->d_gchain);
    if (k->d_flags) cudaFree(k->d_flags);
    k->in_a.release(); k->in_b.release(); k->out_a.release(); k->out_b.release(); k->scratch.release(); k->offs.release();
    k->cin_a.release(); k->cin_b.release(); k->cin_q.release(); k->cin_r.release();
    for (auto& kv : k->cells) { if (kv.second.d_consts) cudaFree(kv.second.d_consts); if (kv.second.d_inc) cudaFree(kv.second.d_inc); if (kv.second.d_mtab) cudaFree(kv.second.d_mtab); }
    if (k->stream) cudaStreamDestroy(k->stream);
    delete k;
}
uint32_t pb200_key_n_bits(const pb200_key* k) { return k ? k->n_bits : 0; }
uint32_t pb200_key_words_in(const pb200_key* k) { return k ? k->words_in : 0; }
uint32_t pb200_key_words_out(const pb200_key* k) { return k ? k->words_out : 0; }
int pb200_key_device(const pb200_key* k) { return k ? k->device : -1; }
int pb200_key_n2(const pb200_key* k, uint64_t* out) {
    if (!k || !out) return PB200_ERR_INVALID_ARG;
    k->n2.to_u64_le(out, k->words_out);
    return PB200_OK;
}
const char* pb200_key_engine(const pb200_key* k) {
    if (!k) return "";
    return use_fast(k) ? k->engine_name.c_str() : "simple64";
}
int pb200_key_set_engine(pb200_key* k, int engine) {
    if (!k || engine < 0 || engine > 3) return PB200_ERR_INVALID_ARG;
    if (engine >= 2 && !k->fast) return PB200_ERR_UNSUPPORTED;
    k->engine = engine;
    if (k->fast) { block28_set_mma(k->fast, engine != 2); k->engine_name = block28_name(k->fast); }
    return PB200_OK;
}
void* pb200_key_stream(const pb200_key* k) { return k ? (void*)k->stream : nullptr; }
int pb200_key_chain_counts(const pb200_key* k, uint64_t* n_sqr, uint64_t* n_mul) {
    if (!k || !n_sqr || !n_mul) return PB200_ERR_INVALID_ARG;
    if (use_fast(k)) { block28_chain_counts(k->fast, n_sqr, n_mul); return PB200_OK; }
    // simple64 runs the reference's LSB-first chain: bits(n) squarings; popcount(n) + ~popcount(m) + 1 multiplications
    uint64_t pn = 0; for (size_t i = 0; i < k->n.w.size(); i++) pn += (uint64_t)__builtin_popcount(k->n.w[i]);
    *n_sqr = k->n.bits(); *n_mul = pn + k->n_bits / 2 + 1;
    return PB200_OK;
}
int pb200_key_sync(pb200_key* k) {
    if (!k) return PB200_ERR_INVALID_ARG;
    CU(cudaSetDevice(k->device));
    CU(cudaStreamSynchronize(k->stream));
    return PB200_OK;
}

// reads and clears the device-side range flag
static int take_flags(pb200_key* k) {
    int f = 0;
    CU(cudaMemcpyAsync(&f, k->d_flags, sizeof(int), cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    if (f) { CU(cudaMemsetAsync(k->d_flags, 0, sizeof(int), k->stream)); return PB200_ERR_RANGE; }
    return PB200_OK;
}

// per-key g-chain records: by the witness engine when it serves this key (it builds its table in the same pass), else by
// simple64 (one thread, the reference's chain)
static int ensure_gchain(pb200_key* k) {
    if (k->d_gchain) return PB200_OK;
    CU(cudaMalloc(&k->d_gchain, (size_t)k->n_bits * 2 * k->words_out * sizeof(u64)));
    if (use_fast(k) && block28_witness_supported(k->fast)) CU(block28_witness_prepare(k->fast, k->d_gchain, false, k->stream));
    else CU(simple_gchain(k->d_simple, k->d_gchain, (int)k->n_bits, k->stream));
    return PB200_OK;
}

// witness producer for this key: the block28t witness engine when the fast engine is selected and n fills its
// declared width, else simple64 (both on the GPU).  Prepares the per-key tables on first use.
static int witness_engine(pb200_key* k, bool* fast) {
    *fast = false;
    int rc = ensure_gchain(k); if (rc) return rc;
    if (!use_fast(k) || !block28_witness_supported(k->fast)) return PB200_OK;
    CU(block28_witness_prepare(k->fast, k->d_gchain, true, k->stream));
    *fast = true;
    return PB200_OK;
}
static int run_witness(pb200_key* k, const u64* d_m, const u64* d_r, size_t count, u64* d_c, u64* d_records,
                       const u64* d_offsets, u64* d_digest) {
    bool fast = false;
    int rc = witness_engine(k, &fast); if (rc) return rc;
    if (fast) CU(block28_witness(k->fast, d_m, d_r, count, d_c, d_records, d_offsets, d_digest, k->stream));
    else CU(simple_encrypt(k->d_simple, k->d_gchain, d_m, d_r, count, d_c, d_records, d_offsets, d_digest, k->d_flags, k->stream));
    return PB200_OK;
}

// ---- encrypt --------------------------------------------------------------------------------
int pb200_encrypt_batch_dev(pb200_key* k, const uint64_t* d_m, const uint64_t* d_r, size_t count, uint64_t* d_c) {
    if (!k || (count && (!d_m || !d_r || !d_c))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    if (use_fast(k)) { CU(block28_encrypt(k->fast, (const u64*)d_m, (const u64*)d_r, count, (u64*)d_c, k->stream)); return PB200_OK; }
    int rc = ensure_gchain(k); if (rc) return rc;
    CU(simple_encrypt(k->d_simple, k->d_gchain, (const u64*)d_m, (const u64*)d_r, count, (u64*)d_c, nullptr, nullptr,
                      nullptr, k->d_flags, k->stream));
    return PB200_OK;
}

static int check_inputs(const pb200_key* k, const uint64_t* v, size_t count) {
    for (size_t u = 0; u < count; u++) if (!fits(v + u * k->words_in, k->words_in, k->n_bits)) return PB200_ERR_RANGE;
    return PB200_OK;
}

int pb200_encrypt_batch(pb200_key* k, const uint64_t* m, const uint64_t* r, size_t count, uint64_t* c_out) {
    if (!k || (count && (!m || !r || !c_out))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    if (k->n_bits % 64) { int rc = check_inputs(k, m, count); if (rc) return rc; rc = check_inputs(k, r, count); if (rc) return rc; }
    CU(cudaSetDevice(k->device));
    size_t bin = count * k->words_in * sizeof(u64), bout = count * k->words_out * sizeof(u64);
    CU(k->in_a.reserve(bin)); CU(k->in_b.reserve(bin)); CU(k->out_a.reserve(bout));
    CU(cudaMemcpyAsync(k->in_a.p, m, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cudaMemcpyAsync(k->in_b.p, r, bin, cudaMemcpyHostToDevice, k->stream));
    int rc = pb200_encrypt_batch_dev(k, (const uint64_t*)k->in_a.p, (const uint64_t*)k->in_b.p, count, (uint64_t*)k->out_a.p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c_out, k->out_a.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return use_fast(k) ? PB200_OK : take_flags(k);
}

// ---- add ------------------------------------------------------------------------------------
int pb200_add_batch_dev(pb200_key* k, const uint64_t* d_c1, const uint64_t* d_c2, uint32_t c_words, size_t count,
                        uint64_t* d_out, uint64_t* d_q) {
    if (!k || (count && (!d_c1 || !d_c2 || !d_out)) || c_words == 0 || c_words > k->words_out) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    bool fast = false;
    int rc = witness_engine(k, &fast); if (rc) return rc;
    if (fast) CU(block28_add(k->fast, (const u64*)d_c1, (const u64*)d_c2, (int)c_words, count, (u64*)d_out, (u64*)d_q, k->d_flags, k->stream));
    else CU(simple_add(k->d_simple, (const u64*)d_c1, (const u64*)d_c2, (int)c_words, count, (u64*)d_out, (u64*)d_q, k->d_flags, k->stream));
    return PB200_OK;
}
int pb200_add_batch(pb200_key* k, const uint64_t* c1, const uint64_t* c2, uint32_t c_words, size_t count, uint64_t* out,
                    uint64_t* q_out) {
    if (!k || (count && (!c1 || !c2 || !out)) || c_words == 0 || c_words > k->words_out) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    size_t bin = count * c_words * sizeof(u64), bout = count * k->words_out * sizeof(u64);
    CU(k->in_a.reserve(bin)); CU(k->in_b.reserve(bin)); CU(k->out_a.reserve(bout));
    if (q_out) CU(k->out_b.reserve(bout));
    CU(cudaMemcpyAsync(k->in_a.p, c1, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cudaMemcpyAsync(k->in_b.p, c2, bin, cudaMemcpyHostToDevice, k->stream));
    int rc = pb200_add_batch_dev(k, (const uint64_t*)k->in_a.p, (const uint64_t*)k->in_b.p, c_words, count,
                                 (uint64_t*)k->out_a.p, q_out ? (uint64_t*)k->out_b.p : nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, k->out_a.p, bout, cudaMemcpyDeviceToHost, k->stream));
    if (q_out) CU(cudaMemcpyAsync(q_out, k->out_b.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return take_flags(k);
}

// ---- tally ----------------------------------------------------------------------------------
int pb200_tally_dev(pb200_key* k, const uint64_t* d_c, size_t count, uint64_t* d_out) {
    if (!k || !d_out || (count && !d_c)) return PB200_ERR_INVALID_ARG;
    CU(cudaSetDevice(k->device));
    if (use_fast(k) && count) { CU(block28_tally(k->fast, (const u64*)d_c, count, (u64*)d_out, k->stream)); return PB200_OK; }
    CU(k->scratch.reserve(simple_tally_scratch_words((int)k->words_out) * sizeof(u64)));
    CU(simple_tally(k->d_simple, (int)k->words_out, (const u64*)d_c, count, (u64*)d_out, (u64*)k->scratch.p, k->d_flags, k->stream));
    return PB200_OK;
}
int pb200_tally(pb200_key* k, const uint64_t* c, size_t count, uint64_t* out) {
    if (!k || !out || (count && !c)) return PB200_ERR_INVALID_ARG;
    CU(cudaSetDevice(k->device));
    size_t bin = count * k->words_out * sizeof(u64), bout = k->words_out * sizeof(u64);
    CU(k->in_a.reserve(bin ? bin : 8)); CU(k->out_b.reserve(bout));
    if (bin) CU(cudaMemcpyAsync(k->in_a.p, c, bin, cudaMemcpyHostToDevice, k->stream));
    int rc = pb200_tally_dev(k, (const uint64_t*)k->in_a.p, count, (uint64_t*)k->out_b.p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, k->out_b.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return take_flags(k);
}
int pb200_tally_combine(pb200_key* k, const uint64_t* partials, size_t n_partials, uint64_t* out) {
    // the combine of G <= 8 shard partials is the same fold; use the exact simple engine for it
    if (!k || !out || (n_partials && !partials)) return PB200_ERR_INVALID_ARG;
    int saved = k->engine; k->engine = 1;
    int rc = pb200_tally(k, partials, n_partials, out);
    k->engine = saved;
    return rc;
}

// ---- witness --------------------------------------------------------------------------------
static uint64_t popcount_words(const uint64_t* v, uint32_t words) {
    uint64_t c = 0; for (uint32_t i = 0; i < words; i++) c += (uint64_t)__builtin_popcountll(v[i]); return c;
}
uint64_t pb200_witness_records_for(const pb200_key* k, const uint64_t* m) {
    if (!k || !m) return 0;
    uint64_t pn = 0; for (size_t i = 0; i < k->n.w.size(); i++) pn += (uint64_t)__builtin_popcount(k->n.w[i]);
    return popcount_words(m, k->words_in) + k->n.bits() + pn + 1;
}

int pb200_key_g_chain(pb200_key* k, uint64_t* records_out) {
    if (!k || !records_out) return PB200_ERR_INVALID_ARG;
    CU(cudaSetDevice(k->device));
    int rc = ensure_gchain(k); if (rc) return rc;
    CU(cudaMemcpyAsync(records_out, k->d_gchain, (size_t)k->n_bits * 2 * k->words_out * sizeof(u64), cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return take_flags(k);
}

int pb200_encrypt_witness_batch(pb200_key* k, const uint64_t* m, const uint64_t* r, size_t count, uint64_t* c_out,
                                size_t max_chunk_units, pb200_witness_sink_fn sink, void* user) {
    if (!k || !sink || (count && (!m || !r))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    int rc = check_inputs(k, m, count); if (rc) return rc;
    rc = check_inputs(k, r, count); if (rc) return rc;
    CU(cudaSetDevice(k->device));
    const size_t rec_words = 2 * (size_t)k->words_out;
    uint64_t fixed = pb200_witness_records_for(k, m) - popcount_words(m, k->words_in);  // bits(n)+popcount(n)+1
    size_t max_unit_records = (size_t)fixed + k->n_bits;
    // bytes of staged records per chunk: a chunk costs one kernel launch whose latency is that of ONE unit's chain (147 ms at
    // |n| = 2048) however few units it holds, so chunks are made as large as a 2 GiB staging buffer allows when memory is plentiful
    size_t budget = (size_t)256 << 20;
    { size_t fr = 0, tot = 0; if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr > ((size_t)16 << 30)) budget = (size_t)2 << 30; }
    size_t auto_units = budget / (max_unit_records * rec_words * sizeof(u64));
    if (auto_units < 1) auto_units = 1;
    size_t chunk_units = max_chunk_units ? max_chunk_units : auto_units;
    std::vector<uint64_t> offsets, host_records;
    std::vector<uint32_t> gcounts;
    for (size_t first = 0; first < count; first += chunk_units) {
        size_t nu = count - first < chunk_units ? count - first : chunk_units;
        offsets.assign(nu + 1, 0); gcounts.assign(nu, 0);
        for (size_t u = 0; u < nu; u++) {
            uint64_t pc = popcount_words(m + (first + u) * k->words_in, k->words_in);
            gcounts[u] = (uint32_t)pc;
            offsets[u + 1] = offsets[u] + pc + fixed;
        }
        size_t total_records = (size_t)offsets[nu];
        size_t bin = nu * k->words_in * sizeof(u64), bout = nu * k->words_out * sizeof(u64);
        CU(k->in_a.reserve(bin)); CU(k->in_b.reserve(bin)); CU(k->out_a.reserve(bout));
        CU(k->offs.reserve((nu + 1) * sizeof(u64)));
        CU(k->scratch.reserve(total_records * rec_words * sizeof(u64)));
        CU(cudaMemcpyAsync(k->in_a.p, m + first * k->words_in, bin, cudaMemcpyHostToDevice, k->stream));
        CU(cudaMemcpyAsync(k->in_b.p, r + first * k->words_in, bin, cudaMemcpyHostToDevice, k->stream));
        CU(cudaMemcpyAsync(k->offs.p, offsets.data(), (nu + 1) * sizeof(u64), cudaMemcpyHostToDevice, k->stream));
        rc = run_witness(k, (const u64*)k->in_a.p, (const u64*)k->in_b.p, nu, (u64*)k->out_a.p, (u64*)k->scratch.p,
                         (const u64*)k->offs.p, nullptr);
        if (rc) return rc;
        host_records.resize(total_records * rec_words);
        CU(cudaMemcpyAsync(host_records.data(), k->scratch.p, total_records * rec_words * sizeof(u64), cudaMemcpyDeviceToHost, k->stream));
        if (c_out) CU(cudaMemcpyAsync(c_out + first * k->words_out, k->out_a.p, bout, cudaMemcpyDeviceToHost, k->stream));
        CU(cudaStreamSynchronize(k->stream));
        rc = take_flags(k); if (rc) return rc;
        pb200_witness_chunk ch;
        ch.first_unit = first; ch.n_units = nu; ch.words_out = k->words_out;
        ch.offsets = offsets.data(); ch.records = host_records.data(); ch.g_mul_counts = gcounts.data();
        if (sink(user, &ch) != 0) return PB200_ERR_SINK;
    }
    return PB200_OK;
}

int pb200_encrypt_witness_digest(pb200_key* k, const uint64_t* m, const uint64_t* r, size_t count, uint64_t* c_out,
                                 uint64_t* digest_out) {
    if (!k || !digest_out || (count && (!m || !r))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    int rc = check_inputs(k, m, count); if (rc) return rc;
    rc = check_inputs(k, r, count); if (rc) return rc;
    CU(cudaSetDevice(k->device));
    size_t bin = count * k->words_in * sizeof(u64), bout = count * k->words_out * sizeof(u64);
    CU(k->in_a.reserve(bin)); CU(k->in_b.reserve(bin)); CU(k->out_a.reserve(bout)); CU(k->out_b.reserve(count * sizeof(u64)));
    CU(cudaMemcpyAsync(k->in_a.p, m, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cudaMemcpyAsync(k->in_b.p, r, bin, cudaMemcpyHostToDevice, k->stream));
    rc = run_witness(k, (const u64*)k->in_a.p, (const u64*)k->in_b.p, count, (u64*)k->out_a.p, nullptr, nullptr, (u64*)k->out_b.p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(digest_out, k->out_b.p, count * sizeof(u64), cudaMemcpyDeviceToHost, k->stream));
    if (c_out) CU(cudaMemcpyAsync(c_out, k->out_a.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return take_flags(k);
}

int pb200_encrypt_witness_digest_dev(pb200_key* k, const uint64_t* d_m, const uint64_t* d_r, size_t count, uint64_t* d_c,
                                     uint64_t* d_digest) {
    if (!k || !d_digest || (count && (!d_m || !d_r))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    return run_witness(k, (const u64*)d_m, (const u64*)d_r, count, (u64*)d_c, nullptr, nullptr, (u64*)d_digest);
}
const char* pb200_key_witness_engine(pb200_key* k) {
    if (!k) return "";
    return (use_fast(k) && block28_witness_supported(k->fast)) ? "block28w" : "simple64";
}

// ---- K4: advice cells ------------------------------------------------------------------------------------
static BigInt low_bits(const BigInt& v, size_t bits) { return BigInt::sub(v, BigInt::shl(BigInt::shr(v, bits), bits)); }
static void put128(std::vector<u64>& dst, const BigInt& v) { u64 w[2]; v.to_u64_le(w, 2); dst.push_back(w[0]); dst.push_back(w[1]); }

static int cell_ctx(pb200_key* k, uint32_t lookup_bits, pb200_key::CellCtx** out) {
    if (lookup_bits > 32) return PB200_ERR_INVALID_ARG;
    auto it = k->cells.find(lookup_bits);
    if (it != k->cells.end()) { *out = &it->second; return PB200_OK; }
    pb200_key::CellCtx C;
    CellLayout& Y = C.Y;
    const int lb = (int)k->limb_bits, L = (int)(2 * k->n_bits / k->limb_bits), NC = 2 * L - 1, kn = (int)(k->n_bits / k->limb_bits);
    Y.L = L; Y.limb_bits = lb; Y.lookup_bits = (int)lookup_bits;
    BigInt B1 = BigInt::sub(BigInt::pow2(lb), BigInt(1));
    BigInt word_max = BigInt::add(BigInt::mul(BigInt((uint64_t)L), BigInt::mul(B1, B1)), B1);       // A.6
    Y.carry_bits = (int)BigInt::shl(word_max, 1).bits() - lb;
    Y.kl = lookup_bits ? (lb + (int)lookup_bits - 1) / (int)lookup_bits : 0;
    Y.xl = lookup_bits && (lb % (int)lookup_bits) ? 1 : 0;
    Y.cpl = 1 + Y.kl + Y.xl;
    Y.kc = lookup_bits ? (Y.carry_bits + (int)lookup_bits - 1) / (int)lookup_bits : 0;
    Y.xc = lookup_bits && (Y.carry_bits % (int)lookup_bits) ? 1 : 0;
    Y.off_rem = L * Y.cpl; Y.off_ab = 2 * L * Y.cpl; Y.off_qn = Y.off_ab + NC; Y.off_qnp = Y.off_qn + NC; Y.off_eq = Y.off_qnp + NC;
    Y.eq_stride = 4 + Y.kc + Y.xc;
    Y.n_cells = Y.off_eq + (NC - 1) * Y.eq_stride + 4 + 1;
    std::vector<u64> h;
    for (int i = 0; i < L; i++) put128(h, low_bits(BigInt::shr(k->n2, (size_t)i * lb), lb));
    { u64 w[4]; word_max.to_u64_le(w, 4); h.insert(h.end(), w, w + 4); }
    std::vector<u64> qa, ma;
    BigInt acc;
    for (int i = 0; i < NC; i++) {
        acc = BigInt::add(acc, word_max);
        BigInt qacc = BigInt::shr(acc, lb), macc = low_bits(acc, lb);
        put128(qa, qacc); put128(ma, macc);
        acc = qacc;
    }
    h.insert(h.end(), qa.begin(), qa.end()); h.insert(h.end(), ma.begin(), ma.end());
    {   // Montgomery forms (x * 2^256 mod p, BN254 Fr) of the per-key q_acc / mod_acc cells; 32-byte aligned in the buffer
        static const uint64_t FRP[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
        BigInt P = BigInt::from_u64_le(FRP, 4);
        for (int which = 0; which < 2; which++)
            for (int i = 0; i < NC; i++) {
                const std::vector<u64>& src = which ? ma : qa;
                u64 w2[2] = {src[2 * i], src[2 * i + 1]};
                BigInt m = BigInt::mod(BigInt::shl(BigInt::from_u64_le(w2, 2), 256), P);
                u64 w4[4]; m.to_u64_le(w4, 4);
                h.insert(h.end(), w4, w4 + 4);
            }
    }
    // RefreshAux::new(limb_bits, kn, kn) (A.3): how far each column of n*n can spill
    std::vector<BigInt> vals(2 * kn - 1);
    for (int i = 0; i < kn; i++) for (int j = 0; j < kn; j++) vals[i + j] = BigInt::add(vals[i + j], BigInt::mul(B1, B1));
    std::vector<int> inc;
    for (size_t i = 0; i < vals.size(); i++) {
        BigInt v = vals[i], carry = BigInt::shr(v, lb);
        int cnt = 0; size_t kk = 1;
        while (!carry.is_zero()) {
            if (i + kk >= vals.size()) vals.push_back(BigInt());
            vals[i + kk] = BigInt::add(vals[i + kk], low_bits(carry, lb));
            carry = BigInt::shr(carry, lb); cnt++; kk++;
        }
        vals[i] = low_bits(v, lb);
        inc.push_back(cnt);
    }
    C.n_out = (int)inc.size();
    C.n2_cells = (2 * kn - 1) + (lookup_bits ? C.n_out * (Y.kl + Y.xl) : 0);
    for (int i = 0; i < 2 * kn - 1; i++) C.n2_cells += 2 * (inc[i] + 1);
    CU(cudaMalloc(&C.d_consts, h.size() * sizeof(u64)));
    CU(cudaMemcpy(C.d_consts, h.data(), h.size() * sizeof(u64), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&C.d_inc, inc.size() * sizeof(int)));
    CU(cudaMemcpy(C.d_inc, inc.data(), inc.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (lookup_bits >= 1 && lookup_bits <= 16) {       // chunk-value table for the Montgomery output (2 MB at 16 bits)
        CU(cudaMalloc(&C.d_mtab, ((size_t)32) << lookup_bits));
        CU(cells_mont_table((int)lookup_bits, C.d_mtab, k->stream));
        CU(cudaStreamSynchronize(k->stream));
    }
    k->cells[lookup_bits] = C;
    *out = &k->cells[lookup_bits];
    return PB200_OK;
}

int pb200_cells_layout(pb200_key* k, uint32_t lookup_bits, pb200_cell_layout* out) {
    if (!k || !out) return PB200_ERR_INVALID_ARG;
    CU(cudaSetDevice(k->device));
    pb200_key::CellCtx* C = nullptr;
    int rc = cell_ctx(k, lookup_bits, &C); if (rc) return rc;
    out->limbs = (uint32_t)C->Y.L; out->cells_per_limb = (uint32_t)C->Y.cpl; out->carry_bits = (uint32_t)C->Y.carry_bits;
    out->cells_per_mulmod = (uint32_t)C->Y.n_cells; out->cells_n2 = (uint32_t)C->n2_cells;
    out->off_rem = (uint32_t)C->Y.off_rem; out->off_ab = (uint32_t)C->Y.off_ab; out->off_qn = (uint32_t)C->Y.off_qn;
    out->off_qn_rem = (uint32_t)C->Y.off_qnp; out->off_eq = (uint32_t)C->Y.off_eq; out->eq_stride = (uint32_t)C->Y.eq_stride;
    return PB200_OK;
}

int pb200_mulmod_cells_batch_dev(pb200_key* k, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_q, const uint64_t* d_rem,
                                 size_t count, uint32_t lookup_bits, int montgomery, uint64_t* d_cells) {
    if (!k || (count && (!d_a || !d_b || !d_q || !d_rem || !d_cells))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    pb200_key::CellCtx* C = nullptr;
    int rc = cell_ctx(k, lookup_bits, &C); if (rc) return rc;
    CU(cells_mulmod(C->Y, C->d_consts, (const u64*)d_a, (const u64*)d_b, (const u64*)d_q, (const u64*)d_rem, count, (int)k->words_out,
                    montgomery ? 1 : 0, (u64*)d_cells, k->d_flags, k->sms, C->d_mtab, k->stream));
    return PB200_OK;
}

int pb200_mulmod_cells_batch(pb200_key* k, const uint64_t* a, const uint64_t* b, const uint64_t* q, const uint64_t* rem, size_t count,
                             uint32_t lookup_bits, int montgomery, uint64_t* cells_out) {
    if (!k || (count && (!a || !b || !q || !rem || !cells_out))) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    pb200_key::CellCtx* C = nullptr;
    int rc = cell_ctx(k, lookup_bits, &C); if (rc) return rc;
    const size_t bin = count * k->words_out * sizeof(u64), bout = count * (size_t)C->Y.n_cells * 32;
    CU(k->cin_a.reserve(bin)); CU(k->cin_b.reserve(bin)); CU(k->cin_q.reserve(bin)); CU(k->cin_r.reserve(bin)); CU(k->scratch.reserve(bout));
    CU(cudaMemcpyAsync(k->cin_a.p, a, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cudaMemcpyAsync(k->cin_b.p, b, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cudaMemcpyAsync(k->cin_q.p, q, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cudaMemcpyAsync(k->cin_r.p, rem, bin, cudaMemcpyHostToDevice, k->stream));
    rc = pb200_mulmod_cells_batch_dev(k, (const uint64_t*)k->cin_a.p, (const uint64_t*)k->cin_b.p, (const uint64_t*)k->cin_q.p,
                                      (const uint64_t*)k->cin_r.p, count, lookup_bits, montgomery, (uint64_t*)k->scratch.p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(cells_out, k->scratch.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    int f = 0;
    CU(cudaMemcpy(&f, k->d_flags, sizeof(int), cudaMemcpyDeviceToHost));
    if (f) { CU(cudaMemset(k->d_flags, 0, sizeof(int))); return PB200_ERR_RANGE; }   // (q, rem) do not satisfy a*b = q*n^2 + rem
    return PB200_OK;
}

int pb200_assign_cells_batch(pb200_key* k, const uint64_t* values, size_t count, uint32_t value_bits, uint32_t lookup_bits,
                             int montgomery, uint64_t* cells_out) {
    if (!k || !values || !cells_out || value_bits == 0 || value_bits % k->limb_bits || lookup_bits > 32) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    const int lb = (int)k->limb_bits, nl = (int)(value_bits / k->limb_bits);
    const int kl = lookup_bits ? (lb + (int)lookup_bits - 1) / (int)lookup_bits : 0, xl = lookup_bits && (lb % (int)lookup_bits) ? 1 : 0;
    const int cpl = 1 + kl + xl, words = (int)PB200_WORDS(value_bits);
    const size_t bin = count * words * sizeof(u64), bout = count * (size_t)nl * cpl * 32;
    CU(k->cin_a.reserve(bin)); CU(k->scratch.reserve(bout));
    CU(cudaMemcpyAsync(k->cin_a.p, values, bin, cudaMemcpyHostToDevice, k->stream));
    CU(cells_assign((const u64*)k->cin_a.p, count, words, nl, lb, (int)lookup_bits, kl, cpl, montgomery ? 1 : 0, (u64*)k->scratch.p, k->stream));
    CU(cudaMemcpyAsync(cells_out, k->scratch.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return PB200_OK;
}

int pb200_key_n2_cells(pb200_key* k, uint32_t lookup_bits, int montgomery, uint64_t* cells_out) {
    if (!k || !cells_out) return PB200_ERR_INVALID_ARG;
    CU(cudaSetDevice(k->device));
    pb200_key::CellCtx* C = nullptr;
    int rc = cell_ctx(k, lookup_bits, &C); if (rc) return rc;
    std::vector<u64> nw(k->words_in);
    k->n.to_u64_le(nw.data(), k->words_in);
    const size_t bout = (size_t)C->n2_cells * 32;
    CU(k->cin_a.reserve(nw.size() * sizeof(u64))); CU(k->scratch.reserve(bout)); CU(k->offs.reserve(sizeof(int)));
    CU(cudaMemcpyAsync(k->cin_a.p, nw.data(), nw.size() * sizeof(u64), cudaMemcpyHostToDevice, k->stream));
    CU(cells_n2((const u64*)k->cin_a.p, (int)k->words_in, (int)(k->n_bits / k->limb_bits), (int)k->limb_bits, (int)lookup_bits, C->Y.kl,
                C->Y.xl, C->d_inc, C->n_out, montgomery ? 1 : 0, (u64*)k->scratch.p, (int*)k->offs.p, k->d_flags, k->stream));
    int written = 0;
    CU(cudaMemcpyAsync(&written, k->offs.p, sizeof(int), cudaMemcpyDeviceToHost, k->stream));
    CU(cudaMemcpyAsync(cells_out, k->scratch.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    if (written != C->n2_cells) { t_cuda_error = "n2 cell count mismatch"; return PB200_ERR_CUDA; }
    return take_flags(k);
}

// ---- limb formatting ------------------------------------------------------------------------
int pb200_repack_limbs(pb200_key* k, const uint64_t* values, size_t count, uint32_t value_bits, uint32_t limb_bits,
                       uint64_t* limbs_out) {
    if (!k || !values || !limbs_out || limb_bits == 0 || limb_bits > 128 || value_bits % limb_bits) return PB200_ERR_INVALID_ARG;
    if (!count) return PB200_OK;
    CU(cudaSetDevice(k->device));
    uint32_t wpv = PB200_WORDS(value_bits), nl = value_bits / limb_bits;
    size_t bin = count * wpv * sizeof(u64), bout = count * nl * 2 * sizeof(u64);
    CU(k->in_a.reserve(bin)); CU(k->out_a.reserve(bout));
    CU(cudaMemcpyAsync(k->in_a.p, values, bin, cudaMemcpyHostToDevice, k->stream));
    CU(repack_limbs((const u64*)k->in_a.p, count, (int)wpv, (int)value_bits, (int)limb_bits, (u64*)k->out_a.p, k->stream));
    CU(cudaMemcpyAsync(limbs_out, k->out_a.p, bout, cudaMemcpyDeviceToHost, k->stream));
    CU(cudaStreamSynchronize(k->stream));
    return PB200_OK;
}

}  // extern "C"
