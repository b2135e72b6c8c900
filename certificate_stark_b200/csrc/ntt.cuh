// Number-theoretic transforms over f63 for batches of columns (K1/K2 of SURVEY.md section 2.2): what winterfell's
// fft::interpolate_poly / evaluate_poly_with_offset do for the reference's prover (Prover::prove entered at
// /root/reference/src/lib.rs:140), re-designed for the GPU:
//
//   * every transform is a size-n natural-order -> natural-order NTT of one column (n = trace length or a periodic
//     column's cycle); the size n*blowup LDE is never formed as one transform: the LDE domain offset*<w_lde> is the
//     union of `blowup` cosets  s_k*<w_n>,  s_k = offset * w_lde^k, and coset k is the plain size-n NTT of the
//     coefficients pre-scaled by s_k^m.  LDE row j = k + blowup*i lives at  lde[k][column][i]  ("coset-major").
//   * a transform of n = n1*n2 points is two passes over HBM (four-step): pass A does n2 strided sub-transforms of
//     size n1 from a tile of T adjacent lanes staged in shared memory, multiplies by w_n^(k1*i2) and writes the tile
//     back; pass B does the n1 contiguous sub-transforms of size n2 and writes the transposed result.  n <= 2^11
//     needs a single pass.  Both passes read and write T*8-byte contiguous segments.
#pragma once
#include "dev.cuh"

namespace csg {

// W[j] = w^j for the primitive 2^logn-th root of unity w (Montgomery form), j < 2^logn.  One table serves forward
// and inverse transforms of every size dividing 2^logn, the four-step twiddles and the x-coordinates of the domain.
struct RootTable {
    DBuf<fe> W;
    unsigned logn = 0;
    void build(unsigned logn, Stream &st);
};

struct NttScratch {
    DBuf<fe> tmp;     // pass A output (ncols * n per coset processed together)
    DBuf<fe> scale;   // per-coset pre-scale tables
    DBuf<fe> shifts;  // coset shifts (device copy)
};

// coefficients: out[c*out_stride + m], from evaluations in[c*in_stride + i] over <w_n>; includes the 1/n scaling.
// out_factor != 0: every output is also multiplied by it (R^2 turns a transform of canonical words into Montgomery-form
// coefficients at no cost); raw_input: the words come from a caller and may be anything below 2^64 (brought below 2p on load)
void intt_columns(const RootTable &rt, NttScratch &sc, const fe *in, size_t in_stride, fe *out, size_t out_stride, size_t ncols,
                  unsigned logn, Stream &st, fe out_factor = 0, bool raw_input = false);
// evaluations of each column polynomial over the cosets shift[z]*<w_n>:
//   out[z*out_coset_stride + c*out_col_stride + i] = sum_m coeffs[c*in_stride + m] * shift[z]^m * w_n^(m*i)
void coset_ntt_columns(const RootTable &rt, NttScratch &sc, const fe *coeffs, size_t in_stride, fe *out, size_t out_col_stride,
                       size_t out_coset_stride, size_t ncols, unsigned logn, const fe *shifts_host, size_t ncosets, Stream &st);
// the same with pre-built per-coset scale tables (a proof extends its columns in chunks, overlapped with the H2D copy)
struct CosetTables {
    DBuf<fe> shifts, tables, full;   // full[z][m] = shift_z^m for two-pass sizes
    size_t ncosets = 0;
    unsigned logn = 0;
    bool has_full = false;
    void build(const fe *shifts_host, size_t ncosets, unsigned logn, Stream &st);
};
void coset_ntt_columns(const RootTable &rt, NttScratch &sc, const fe *coeffs, size_t in_stride, fe *out, size_t out_col_stride,
                       size_t out_coset_stride, size_t ncols, unsigned logn, const CosetTables &ct, Stream &st);
void coset_ntt_columns(const RootTable &rt, NttScratch &sc, const fe *coeffs, size_t in_stride, fe *out, size_t out_col_stride,
                       size_t out_coset_stride, size_t ncols, unsigned logn, const fe *tables_dev, size_t ncosets, Stream &st, int in_coset_stride,
                       const fe *full_tables = nullptr);
// independent entries, each with its own coefficients and shift: out[z*n + i] = sum_m in[z*n + m] * shift[z]^m * w_n^(m*i)
void coset_ntt_entries(const RootTable &rt, NttScratch &sc, const fe *in, fe *out, size_t nentries, unsigned logn, const fe *shifts_host, Stream &st);
// inverse of the above for one coset per batch entry z: coefficients of the polynomial whose evaluations over
// shift[z]*<w_n> are in[z*in_stride + i]; out[z*out_stride + m]  (interpolate_poly_with_offset)
void coset_intt_columns(const RootTable &rt, NttScratch &sc, const fe *in, size_t in_stride, fe *out, size_t out_stride,
                        unsigned logn, const fe *shift_inv_host, size_t ncosets, Stream &st);

}  // namespace csg
