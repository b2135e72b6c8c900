// The streaming kernels between the big transforms of Prover::prove (K6-K10 of SURVEY.md section 2.2): representation
// change of the incoming trace, the cross-coset step that turns per-coset interpolants into composition columns, the
// out-of-domain evaluations, the DEEP composition and the FRI degree-respecting projection.  All restate the published
// winterfell v0.3 protocol (the fork at Cargo.toml:20 of the reference is not vendored) as flat data-parallel kernels.
#pragma once
#include "dev.cuh"

namespace csg {

// canonical u64 -> Montgomery (BaseElement::new) and back
// canonical words (or, with already_montgomery, Montgomery words as winterfell's TraceTable stores them) -> reduced Montgomery form
void to_montgomery(const uint64_t *in, fe *out, size_t count, Stream &st, bool already_montgomery = false);

// e[k][m] (k < ce): coefficient m of the size-n interpolant of C on ce coset k, already divided by s_k^m.
// cols[r][q] = coefficient q*ce + r of the degree < ce*n composition polynomial (CompositionPoly::new's transposition).
// mat[t*ce + k] = offset^(-n t) / ce * w_ce^(-k t)
// col_stride: distance between consecutive output columns (0: n)
void composition_columns(const fe *e, fe *cols, size_t n, unsigned ce, const fe *mat_host, Stream &st, size_t col_stride = 0);

// out[p][j'][m] = sum_j mat[j'*L + j] * in[p][j][m] for npolys blocks of L x n elements; out[e] = sum_k in[k*stride + e]
void coset_mix(const fe *in, fe *out, size_t n, unsigned L, size_t npolys, const fe *mat_host, Stream &st);
void sum_slices(const fe *in, fe *out, size_t stride, unsigned count, Stream &st);
// sharded form: input coset j of polynomial p at in[(j / Ll) * rank_stride + (p * Ll + j % Ll) * n] (slices gathered from L / Ll
// ranks), mat_host: the Ll rows of the L x L map that belong to this rank; out[p][Ll][n]
void coset_mix_sharded(const fe *in, fe *out, size_t n, unsigned L, unsigned Ll, size_t rank_stride, size_t npolys, const fe *mat_host, Stream &st);

// values[p * ncols + c] = poly_c(points[p]) for ncols polynomials of n coefficients at polys[c * stride ..]; the final
// reduction over per-CTA partial sums runs on the host (a few hundred KB)
void eval_polys_at(const fe *polys, size_t stride, size_t ncols, size_t n, const fe *points_host, size_t npoints, fe *values_host,
                   DBuf<fe> &scratch, Stream &st);

// out[t][m] = sum_c coef[t][c] * polys[c][m], t < ncomb (linear combinations of coefficient vectors); coef on the host
void combine_polys(const fe *polys, size_t stride, size_t ncols, size_t n, const fe *coef_host, size_t ncomb, fe *out, size_t out_stride,
                   DBuf<fe> &scratch, Stream &st);

struct DeepArgs {
    fe z, zg, zm;         // the three out-of-domain points
    fe az, bzg, czm;      // sum alpha_c T_c(z), sum beta_c T_c(zg), sum delta_r H_r(z^m)
    fe lambda, mu;        // degree adjustment (lambda + mu x)
    fe shift[32];         // coset shifts s_k of the LDE domain
    unsigned ncosets;     // blowup
};
// abc: coset-major LDE of the three combined polynomials, abc[(k*3 + t)*n + i]; W = root table of size n.
// deep[j] for natural LDE index j = k + ncosets*i:
//   ((A - az)/(x - z) + (B - bzg)/(x - zg) + (C - czm)/(x - zm)) * (lambda + mu x)
void deep_quotients(const fe *abc, const fe *W, size_t n, const DeepArgs &a, fe *deep, Stream &st);

struct FoldArgs {
    fe alpha, offset_inv, zeta_inv, quarter;
    fe small[32];       // w_m^-(i_lo), i_lo < 2^extra, where m = 2^logm may exceed the root table size 2^logW
    unsigned logm, logW;
};
// FRI folding factor 4: out[i] = interpolant of {e[i], e[i+q], e[i+2q], e[i+3q]} on x_i*{1, zeta, zeta^2, zeta^3} at alpha
void fri_fold4(const fe *evals, size_t m, const fe *W, const FoldArgs &a, fe *out, Stream &st);

// rows[t][c] (canonical) = element (c, position[t]) of a coset-major matrix; position j = k + ncosets*i
// sub > 1: column c lives at (c / sub) * col_stride + (c % sub) * sub_stride (planes of extension-field elements)
void gather_rows(const fe *data, unsigned width, unsigned ncosets, size_t coset_stride, size_t col_stride, const uint32_t *pos_dev,
                 size_t npos, uint64_t *rows_dev, Stream &st, unsigned sub = 1, size_t sub_stride = 0);
void from_montgomery(const fe *in, uint64_t *out, size_t count, Stream &st);
// coset-major lde[(k*width + c)*n + i] -> canonical natural-order columns out[c*(n*ncosets) + k + ncosets*i]
void coset_major_to_natural(const fe *lde, unsigned width, unsigned ncosets, size_t n, uint64_t *out, Stream &st);

// arithmetic self-test on the device: number of (random and edge-case) inputs on which the fast Montgomery reduction
// disagrees with the textbook one; must be 0
long long redc_selftest(Stream &st);
long long field_selftest(Stream &st);   // field_selfcheck.cuh on 32768 threads: mismatches (0 = pass)

}  // namespace csg
