// cells.hpp — layout of the advice cells of one mul_mod group and the launchers of cells.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pb200 {

// Cell order of one BigUintChip::mul_mod(a, b, n^2) group:
//   [0, off_rem)        q:   per limb  [limb, kl chunks, xl shifted top chunk]        (assign_integer + range_check)
//   [off_rem, off_ab)   rem: same
//   [off_ab, off_qn)    ab columns          (2L-1)
//   [off_qn, off_qnp)   q*n^2 columns       (2L-1)
//   [off_qnp, off_eq)   q*n^2 + rem sums    (2L-1)
//   [off_eq, n_cells-1) per column i: carry_{i+1}, cs_i, q_acc_i, mod_acc_i [, kc chunks, xc shifted top chunk of the carry; i < 2L-2]
//   n_cells-1           eq
struct CellLayout {
    int L, limb_bits, lookup_bits;
    int kl, xl, cpl;            // chunks per limb, extra cell, cells per assigned limb (1 + kl + xl)
    int carry_bits, kc, xc;     // is_equal_muled carries: width, chunks, extra cell
    int off_rem, off_ab, off_qn, off_qnp, off_eq, eq_stride, n_cells;
};

size_t cells_mulmod_smem(const CellLayout& Y);
cudaError_t cells_mulmod(const CellLayout& Y, const uint64_t* d_consts, const uint64_t* d_a, const uint64_t* d_b, const uint64_t* d_q,
                         const uint64_t* d_rem, size_t count, int words, int mont, uint64_t* d_out, int* d_flags, int sms,
                         const uint64_t* d_mtab /* nullable: Montgomery forms of 0 .. 2^lookup_bits-1 */, cudaStream_t st);
cudaError_t cells_mont_table(int lookup_bits, uint64_t* d_tab /* 2^lookup_bits * 4 words */, cudaStream_t st);
cudaError_t cells_assign(const uint64_t* d_vals, size_t count, int words, int nl, int limb_bits, int lookup, int k, int cpl, int mont,
                         uint64_t* d_out, cudaStream_t st);
cudaError_t cells_n2(const uint64_t* d_n_words, int words, int kn, int limb_bits, int lookup, int kl, int xl, const int* d_inc, int n_out,
                     int mont, uint64_t* d_out, int* d_n_written, int* d_flags, cudaStream_t st);

}  // namespace pb200
