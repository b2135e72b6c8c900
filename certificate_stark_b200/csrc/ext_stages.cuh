// Protocol stages over the extension field E (ext.cuh) for FieldExtension::Quadratic / Cubic.  The trace, its LDE and
// every transform stay in the base field; E enters through the random challenges:
//   * the merged constraint column, the composition columns and the DEEP polynomial are E-valued: stored as d planes
//     (component j of every element contiguous), so interpolation, LDE and linear combinations reuse the base kernels;
//   * out-of-domain evaluation at z in E = dot products of coefficient columns with the table of powers of z;
//   * DEEP quotients and FRI folding multiply and invert in E point by point.
#pragma once
#include "dev.cuh"
#include "ext.cuh"
#include "stages.cuh"

namespace csg {
using f63::xe;

// tab[(p*d + j)*n + m] = component j of pts[p]^m, m < n
void ext_power_table(int d, const xe *pts_host, size_t npts, size_t n, fe *tab, Stream &st);
// values[c*nw + v] = sum_m polys[c*stride + m] * wt[v*n + m]   (nw weight vectors of n elements on the device; result on the host)
void dot_columns(const fe *polys, size_t stride, size_t ncols, size_t n, const fe *wt, size_t nw, fe *values_host, DBuf<fe> &scratch, Stream &st);

struct DeepArgsX {
    int d;
    xe z, zg, zm, az, bzg, czm, lambda, mu;
    fe shift[32];
    unsigned ncosets;
    f63::ExtConsts k;
};
// abc: coset-major LDE of the 3 combined polynomials as 3*d base columns, abc[(k*3*d + t*d + j)*n + i];
// deep[j*plane + (k + ncosets*i)] = component j of the DEEP evaluation at LDE row k + ncosets*i
void deep_quotients_ext(const fe *abc, const fe *W, size_t n, const DeepArgsX &a, fe *deep, size_t plane, Stream &st);

struct FoldArgsX { FoldArgs f; xe alpha; int d; };   // f.alpha is unused
// FRI folding factor 4 on an E-valued layer of m elements stored as d planes of stride in_plane; out: d planes of stride out_plane
void fri_fold4_ext(const fe *evals, size_t m, size_t in_plane, const fe *W, const FoldArgsX &a, fe *out, size_t out_plane, Stream &st);
// out[i*d + j] = canonical value of component j of element i (serialisation order of E elements)
void planes_to_canonical(const fe *planes, size_t plane, size_t count, int d, uint64_t *out, Stream &st);

}  // namespace csg
