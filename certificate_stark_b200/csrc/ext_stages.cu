// Kernels of the extension-field protocol stages (FieldExtension::Quadratic / Cubic): out-of-domain evaluation at a point of
// E, DEEP quotients with E-valued numerators and denominators, FRI folding of E-valued layers.  See ext_stages.cuh.
#include <vector>

#include "ext_stages.cuh"

namespace csg {
using namespace f63;

namespace {

constexpr unsigned PW_CHUNK = 64;
struct PowArgs { xe pt[4]; int d; };
// tab[(p*d + j)*n + m] = (pt_p^m)[j]: every thread raises the point to the start of its chunk, then walks the chunk
__global__ void ext_power_table_kernel(PowArgs a, unsigned long long n, fe *__restrict__ tab) {
    const unsigned long long m0 = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * PW_CHUNK;
    if (m0 >= n) return;
    const unsigned p = blockIdx.y;
    const int d = a.d;
    const xe pt = a.pt[p];
    xe cur = x_pow(d, pt, m0);
    for (unsigned k = 0; k < PW_CHUNK && m0 + k < n; k++) {
        for (int j = 0; j < d; j++) tab[((unsigned long long)p * d + j) * n + m0 + k] = cur.c[j];
        cur = x_mul(d, cur, pt);
    }
}

constexpr unsigned DOT_THREADS = 256;
// partial[(c * gridDim.x + s) * NW + v] = sum over the s-th slice of m of polys[c*stride + m] * wt[v*n + m]
template <int NW>
__global__ void __launch_bounds__(DOT_THREADS) dot_columns_kernel(const fe *__restrict__ polys, unsigned long long stride, unsigned long long n,
                                                                  const fe *__restrict__ wt, fe *__restrict__ partial) {
    __shared__ fe red[NW][DOT_THREADS];
    const unsigned t = threadIdx.x, c = blockIdx.y;
    const unsigned long long per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    acc192 s[NW];
    for (unsigned long long m = lo + t; m < hi; m += DOT_THREADS) {
        const fe v = polys[c * stride + m];
#pragma unroll
        for (int k = 0; k < NW; k++) s[k].mac(v, wt[k * n + m]);
    }
#pragma unroll
    for (int k = 0; k < NW; k++) red[k][t] = s[k].reduce();
    __syncthreads();
    for (unsigned h = DOT_THREADS / 2; h > 0; h >>= 1) {
        if (t < h)
#pragma unroll
            for (int k = 0; k < NW; k++) red[k][t] = add(red[k][t], red[k][t + h]);
        __syncthreads();
    }
    if (t < NW) partial[((unsigned long long)c * gridDim.x + blockIdx.x) * NW + t] = red[t][0];
}

// One thread per LDE point x = s_k w^i (a base-field element): the three numerators are E-valued (d planes each), the
// denominators x - z, x - z g, x - z^ce lie in E; one inversion of their product per point.
__global__ void __launch_bounds__(128) deep_quotients_ext_kernel(const fe *__restrict__ abc, const fe *__restrict__ W, unsigned long long n, DeepArgsX a,
                                                                 fe *__restrict__ deep, unsigned long long plane) {
    const unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x, total = n * a.ncosets;
    if (j >= total) return;
    const int d = a.d;
    const unsigned k = (unsigned)(j % a.ncosets);
    const unsigned long long i = j / a.ncosets;
    const fe x = mul(a.shift[k], W[i]);
    const xe xx = x_from(x);
    const xe d1 = x_sub(xx, a.z), d2 = x_sub(xx, a.zg), d3 = x_sub(xx, a.zm);
    const xe d12 = x_mul(d, d1, d2);
    const xe pinv = x_inv(d, x_mul(d, d12, d3), a.k);
    const xe i3 = x_mul(d, pinv, d12), i12 = x_mul(d, pinv, d3);
    const xe i1 = x_mul(d, i12, d2), i2 = x_mul(d, i12, d1);
    const fe *src = abc + (unsigned long long)k * 3 * d * n + i;
    xe na = x_zero(), nb = x_zero(), nc = x_zero();
    for (int c = 0; c < d; c++) { na.c[c] = src[(unsigned long long)c * n]; nb.c[c] = src[(unsigned long long)(d + c) * n]; nc.c[c] = src[(unsigned long long)(2 * d + c) * n]; }
    na = x_sub(na, a.az); nb = x_sub(nb, a.bzg); nc = x_sub(nc, a.czm);
    const xe s = x_add(x_add(x_mul(d, na, i1), x_mul(d, nb, i2)), x_mul(d, nc, i3));
    const xe r = x_mul(d, s, x_add(a.lambda, x_scale(a.mu, x)));
    for (int c = 0; c < d; c++) deep[(unsigned long long)c * plane + j] = r.c[c];
}

__global__ void fri_fold4_ext_kernel(const fe *__restrict__ e, unsigned long long q, unsigned long long in_plane, const fe *__restrict__ W, FoldArgsX a,
                                     fe *__restrict__ out, unsigned long long out_plane) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= q) return;
    const int d = a.d;
    // x_i^-1 = offset^-1 * w_m^-i, from the root table of size 2^logW (w_m^-i = W-entry times a small correction when m > 2^logW)
    const unsigned long long NW = 1ULL << a.f.logW;
    fe winv;
    if (a.f.logm <= a.f.logW) winv = W[(NW - (i << (a.f.logW - a.f.logm))) & (NW - 1)];
    else {
        const unsigned extra = a.f.logm - a.f.logW;
        winv = mul(W[(NW - (i >> extra)) & (NW - 1)], a.f.small[i & ((1u << extra) - 1)]);
    }
    const fe x_inv = mul(a.f.offset_inv, winv);
    xe v[4];
    for (int t = 0; t < 4; t++) { v[t] = x_zero(); for (int c = 0; c < d; c++) v[t].c[c] = e[(unsigned long long)c * in_plane + i + t * q]; }
    const xe s02 = x_add(v[0], v[2]), d02 = x_sub(v[0], v[2]), s13 = x_add(v[1], v[3]), d13 = x_scale(x_sub(v[1], v[3]), a.f.zeta_inv);
    const xe c0 = x_add(s02, s13), c1 = x_add(d02, d13), c2 = x_sub(s02, s13), c3 = x_sub(d02, d13);
    const xe y = x_scale(a.alpha, x_inv);
    xe r = c3;
    r = x_add(x_mul(d, r, y), c2);
    r = x_add(x_mul(d, r, y), c1);
    r = x_add(x_mul(d, r, y), c0);
    r = x_scale(r, a.f.quarter);
    for (int c = 0; c < d; c++) out[(unsigned long long)c * out_plane + i] = r.c[c];
}

// out[i*d + j] = canonical(planes[j*plane + i]): E elements in serialisation order
__global__ void planes_to_canonical_kernel(const fe *__restrict__ in, unsigned long long plane, unsigned long long count, int d, uint64_t *__restrict__ out) {
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (t >= count * d) return;
    out[t] = from_mont(in[(t % d) * plane + t / d]);
}

}  // namespace

void ext_power_table(int d, const xe *pts_host, size_t npts, size_t n, fe *tab, Stream &st) {
    if (npts > 4) throw std::runtime_error("at most 4 points per power table");
    PowArgs a{};
    a.d = d;
    for (size_t p = 0; p < npts; p++) a.pt[p] = pts_host[p];
    const size_t threads = (n + PW_CHUNK - 1) / PW_CHUNK;
    CSG_LAUNCH(st, ext_power_table_kernel, dim3((unsigned)((threads + 127) / 128), (unsigned)npts), 128, 0, a, (unsigned long long)n, tab);
}

void dot_columns(const fe *polys, size_t stride, size_t ncols, size_t n, const fe *wt, size_t nw, fe *values_host, DBuf<fe> &scratch, Stream &st) {
    const unsigned nsplit = n >= (1u << 16) ? 8 : 1;
    scratch.reserve(ncols * nsplit * nw);
    dim3 grid(nsplit, (unsigned)ncols);
#define CSG_DOT(N) CSG_LAUNCH(st, dot_columns_kernel<N>, grid, DOT_THREADS, 0, polys, (unsigned long long)stride, (unsigned long long)n, wt, scratch.p)
    switch (nw) {
    case 1: CSG_DOT(1); break;
    case 2: CSG_DOT(2); break;
    case 3: CSG_DOT(3); break;
    case 4: CSG_DOT(4); break;
    case 6: CSG_DOT(6); break;
    default: throw std::runtime_error("dot_columns: 1, 2, 3, 4 or 6 weight vectors per call");
    }
#undef CSG_DOT
    std::vector<fe> part(ncols * nsplit * nw);
    CSG_CUDA(cudaMemcpyAsync(part.data(), scratch.p, part.size() * sizeof(fe), cudaMemcpyDeviceToHost, st.s));
    CSG_CUDA(cudaStreamSynchronize(st.s));
    for (size_t c = 0; c < ncols; c++)
        for (size_t v = 0; v < nw; v++) {
            fe s = 0;
            for (unsigned b = 0; b < nsplit; b++) s = add(s, part[(c * nsplit + b) * nw + v]);
            values_host[c * nw + v] = s;
        }
}

void deep_quotients_ext(const fe *abc, const fe *W, size_t n, const DeepArgsX &a, fe *deep, size_t plane, Stream &st) {
    const size_t total = n * a.ncosets;
    CSG_LAUNCH(st, deep_quotients_ext_kernel, (unsigned)((total + 127) / 128), 128, 0, abc, W, (unsigned long long)n, a, deep, (unsigned long long)plane);
}

void fri_fold4_ext(const fe *evals, size_t m, size_t in_plane, const fe *W, const FoldArgsX &a, fe *out, size_t out_plane, Stream &st) {
    const size_t q = m / 4;
    CSG_LAUNCH(st, fri_fold4_ext_kernel, (unsigned)((q + 127) / 128), 128, 0, evals, (unsigned long long)q, (unsigned long long)in_plane, W, a, out,
               (unsigned long long)out_plane);
}

void planes_to_canonical(const fe *planes, size_t plane, size_t count, int d, uint64_t *out, Stream &st) {
    CSG_LAUNCH(st, planes_to_canonical_kernel, (unsigned)((count * d + 255) / 256), 256, 0, planes, (unsigned long long)plane, (unsigned long long)count, d, out);
}

}  // namespace csg
