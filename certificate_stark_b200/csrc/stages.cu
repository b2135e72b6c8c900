// See stages.cuh.
#include <vector>

#include "field_selfcheck.cuh"
#include "stages.cuh"

namespace csg {
using namespace f63;

namespace {

// by = R^2: canonical words -> Montgomery form; by = R (the Montgomery form of 1): words that already are in Montgomery form, reduced below p
__global__ void to_mont_kernel(const uint64_t *__restrict__ in, fe *__restrict__ out, unsigned long long count, fe by) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x, step = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < count; i += step) out[i] = mul(in[i], by);
}
__global__ void from_mont_kernel(const fe *__restrict__ in, uint64_t *__restrict__ out, unsigned long long count) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x, step = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < count; i += step) out[i] = from_mont(in[i]);
}
unsigned grid_for(size_t count, unsigned threads) {
    size_t g = (count + threads - 1) / threads;
    const size_t cap = 148 * 16;   // grid-stride kernels: a few waves of the 148 SMs
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

struct CrossMat { fe m[16 * 16]; };
__global__ void composition_columns_kernel(const fe *__restrict__ e, fe *__restrict__ cols, unsigned long long n, unsigned ce, CrossMat M,
                                           unsigned long long col_stride) {
    unsigned long long m = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (m >= n) return;
    fe v[16];
    for (unsigned k = 0; k < ce; k++) v[k] = e[k * n + m];
    const unsigned long long r = m % ce, q0 = m / ce, qs = n / ce;
    for (unsigned t = 0; t < ce; t++) {
        acc192 s;
        for (unsigned k = 0; k < ce; k++) s.mac(M.m[t * ce + k], v[k]);
        cols[r * col_stride + q0 + qs * t] = s.reduce();
    }
}

// out[p][j'][m] = sum_j M[j'*L + j] * in[p][j][m]: the same L x L linear map applied across the L cosets of every polynomial
__global__ void coset_mix_kernel(const fe *__restrict__ in, fe *__restrict__ out, unsigned long long n, unsigned L, CrossMat M) {
    unsigned long long m = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (m >= n) return;
    const fe *src = in + (unsigned long long)blockIdx.y * L * n + m;
    fe *dst = out + (unsigned long long)blockIdx.y * L * n + m;
    fe v[16];
    for (unsigned j = 0; j < L; j++) v[j] = src[j * n];
    for (unsigned t = 0; t < L; t++) {
        acc192 s;
        for (unsigned j = 0; j < L; j++) s.mac(M.m[t * L + j], v[j]);
        dst[t * n] = s.reduce();
    }
}
// the same for a sharded proof: the L input cosets of polynomial p come from L / Ll ranks (coset j at
// in[(j / Ll) * rank_stride + (p * Ll + j % Ll) * n]); only the Ll output cosets of this rank are formed (M: Ll x L), out[p][Ll][n]
__global__ void coset_mix_sharded_kernel(const fe *__restrict__ in, fe *__restrict__ out, unsigned long long n, unsigned L, unsigned Ll,
                                         unsigned long long rank_stride, CrossMat M) {
    unsigned long long m = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (m >= n) return;
    const unsigned p = blockIdx.y;
    fe v[16];
    for (unsigned j = 0; j < L; j++) v[j] = in[(j / Ll) * rank_stride + ((unsigned long long)p * Ll + j % Ll) * n + m];
    fe *dst = out + (unsigned long long)p * Ll * n + m;
    for (unsigned t = 0; t < Ll; t++) {
        acc192 s;
        for (unsigned j = 0; j < L; j++) s.mac(M.m[t * L + j], v[j]);
        dst[t * n] = s.reduce();
    }
}
// out[e] = sum_k in[k * stride + e], k < count
__global__ void sum_slices_kernel(const fe *__restrict__ in, fe *__restrict__ out, unsigned long long stride, unsigned count) {
    unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (e >= stride) return;
    fe s = in[e];
    for (unsigned k = 1; k < count; k++) s = add(s, in[k * stride + e]);
    out[e] = s;
}

constexpr unsigned EVAL_THREADS = 256, EVAL_CHUNK = 32;
struct EvalPoints { fe z[4], z_stride[4]; };   // z_stride = z^EVAL_THREADS
// One pass over a column for ALL points: thread t of a CTA owns the coefficients base + t + j*THREADS and forms
// sum_j c_j (z^THREADS)^j per point as independent lazily reduced multiply-accumulates against a shared table of powers (the Horner
// form it replaces was one dependent multiply-add chain per thread and point, and read the column once per point).
template <int NPTS>
__global__ void __launch_bounds__(EVAL_THREADS) eval_polys_kernel(const fe *__restrict__ polys, unsigned long long stride, unsigned long long n,
                                                                 EvalPoints pts, fe *__restrict__ partial) {
    __shared__ fe red[EVAL_THREADS];
    __shared__ fe zpow[NPTS][EVAL_CHUNK];
    const unsigned t = threadIdx.x, c = blockIdx.y;
    const unsigned long long base = blockIdx.x * (unsigned long long)(EVAL_THREADS * EVAL_CHUNK);
    const fe *poly = polys + c * stride;
    if (t < NPTS * EVAL_CHUNK) zpow[t / EVAL_CHUNK][t % EVAL_CHUNK] = f63::pow(pts.z_stride[t / EVAL_CHUNK], t % EVAL_CHUNK);
    __syncthreads();
    acc192 v[NPTS];
#pragma unroll 8
    for (int j = 0; j < (int)EVAL_CHUNK; j++) {
        const unsigned long long m = base + t + (unsigned long long)j * EVAL_THREADS;
        const fe cm = m < n ? poly[m] : 0;
#pragma unroll
        for (int p = 0; p < NPTS; p++) v[p].mac(cm, zpow[p][j]);
    }
#pragma unroll
    for (int p = 0; p < NPTS; p++) {
        red[t] = mul(v[p].reduce(), f63::pow(pts.z[p], base + t));
        __syncthreads();
        for (unsigned s = EVAL_THREADS / 2; s > 0; s >>= 1) {
            if (t < s) red[t] = add(red[t], red[t + s]);
            __syncthreads();
        }
        if (t == 0) partial[((unsigned long long)p * gridDim.y + c) * gridDim.x + blockIdx.x] = red[0];
        __syncthreads();
    }
}

template <int NCOMB>
__global__ void combine_polys_kernel(const fe *__restrict__ polys, unsigned long long stride, unsigned ncols, unsigned long long n,
                                     const fe *__restrict__ coef, fe *__restrict__ out, unsigned long long out_stride) {
    unsigned long long m = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (m >= n) return;
    acc192 s[NCOMB];
    for (unsigned c = 0; c < ncols; c++) {
        fe v = polys[c * stride + m];
#pragma unroll
        for (int t = 0; t < NCOMB; t++) s[t].mac(coef[t * ncols + c], v);
    }
#pragma unroll
    for (int t = 0; t < NCOMB; t++) out[t * out_stride + m] = s[t].reduce();
}

// Each thread handles DEEP_CHUNK consecutive-stride elements and inverts the product of all their denominators once
// (Montgomery's trick): 3 multiplications per denominator instead of a ~90-multiplication inversion per element.
constexpr int DEEP_CHUNK = 8, DEEP_THREADS = 128;
__global__ void __launch_bounds__(DEEP_THREADS) deep_quotients_kernel(const fe *__restrict__ abc, const fe *__restrict__ W, unsigned long long n, DeepArgs a,
                                                                     fe *__restrict__ deep) {
    const unsigned long long total = n * a.ncosets;
    const unsigned long long base = blockIdx.x * (unsigned long long)(DEEP_CHUNK * DEEP_THREADS) + threadIdx.x;
    fe den[DEEP_CHUNK], pre[DEEP_CHUNK], xs[DEEP_CHUNK];
    fe acc = ONE;
#pragma unroll
    for (int c = 0; c < DEEP_CHUNK; c++) {
        const unsigned long long j = base + (unsigned long long)c * DEEP_THREADS;
        fe d = ONE, x = 0;
        if (j < total) {
            x = mul(a.shift[j % a.ncosets], W[j / a.ncosets]);
            d = mul(mul(sub(x, a.z), sub(x, a.zg)), sub(x, a.zm));
        }
        xs[c] = x; den[c] = d; pre[c] = acc;
        acc = mul(acc, d);
    }
    fe ainv = inv(acc);
#pragma unroll
    for (int c = DEEP_CHUNK - 1; c >= 0; c--) {
        const unsigned long long j = base + (unsigned long long)c * DEEP_THREADS;
        const fe pinv = mul(ainv, pre[c]);   // 1 / (d1 d2 d3) of element c
        ainv = mul(ainv, den[c]);
        if (j >= total) continue;
        const unsigned k = (unsigned)(j % a.ncosets);
        const unsigned long long i = j / a.ncosets;
        const fe x = xs[c];
        const fe *src = abc + (unsigned long long)k * 3 * n + i;
        const fe na = sub(src[0], a.az), nb = sub(src[n], a.bzg), nc = sub(src[2 * n], a.czm);
        const fe d1 = sub(x, a.z), d2 = sub(x, a.zg), d3 = sub(x, a.zm);
        const fe d12 = mul(d1, d2);
        const fe i3 = mul(pinv, d12), i12 = mul(pinv, d3);   // 1/d3, 1/(d1 d2)
        const fe i1 = mul(i12, d2), i2 = mul(i12, d1);
        fe s = add(add(mul(na, i1), mul(nb, i2)), mul(nc, i3));
        deep[j] = mul(s, add(a.lambda, mul(a.mu, x)));
    }
}

__global__ void fri_fold4_kernel(const fe *__restrict__ e, unsigned long long q, const fe *__restrict__ W, FoldArgs a, fe *__restrict__ out) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= q) return;
    const unsigned long long N = 1ULL << a.logW;
    fe wi;   // w_m^-i
    if (a.logm <= a.logW) wi = W[(N - (i << (a.logW - a.logm))) & (N - 1)];
    else {
        const unsigned extra = a.logm - a.logW;
        wi = mul(a.small[i & ((1u << extra) - 1)], W[(N - (i >> extra)) & (N - 1)]);
    }
    const fe xinv = mul(a.offset_inv, wi);
    const fe v0 = e[i], v1 = e[i + q], v2 = e[i + 2 * q], v3 = e[i + 3 * q];
    const fe s02 = add(v0, v2), d02 = sub(v0, v2), s13 = add(v1, v3), d13 = mul(sub(v1, v3), a.zeta_inv);
    const fe c0 = add(s02, s13), c1 = add(d02, d13), c2 = sub(s02, s13), c3 = sub(d02, d13);
    const fe y = mul(a.alpha, xinv);
    fe r = c3;
    r = add(mul(r, y), c2);
    r = add(mul(r, y), c1);
    r = add(mul(r, y), c0);
    out[i] = mul(r, a.quarter);
}

__global__ void gather_rows_kernel(const fe *__restrict__ data, unsigned width, unsigned ncosets, unsigned long long coset_stride,
                                   unsigned long long col_stride, const uint32_t *__restrict__ pos, unsigned npos, uint64_t *__restrict__ rows,
                                   unsigned sub, unsigned long long sub_stride) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npos * width) return;
    unsigned r = t / width, c = t % width;
    if (pos[r] == 0xFFFFFFFFu) { rows[t] = 0; return; }   // a row another context of a sharded proof owns
    unsigned long long j = pos[r], k = j % ncosets, i = j / ncosets;
    rows[t] = from_mont(data[k * coset_stride + (c / sub) * col_stride + (c % sub) * sub_stride + i]);
}

__global__ void coset_to_natural_kernel(const fe *__restrict__ lde, unsigned width, unsigned ncosets, unsigned long long n, uint64_t *__restrict__ out) {
    unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned c = blockIdx.y;
    if (j >= n * ncosets) return;
    unsigned long long k = j % ncosets, i = j / ncosets;
    out[c * n * ncosets + j] = from_mont(lde[(k * width + c) * n + i]);
}

__global__ void redc_selftest_kernel(unsigned long long *bad, unsigned long long seed) {
    unsigned long long x = seed + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ULL;
    unsigned long long nb = 0;
    for (int it = 0; it < 4096; it++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long lo = x;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long hi = x % P;
        if (it & 1) { lo = (it & 2) ? 0 : (lo << 32); }
        if ((it & 12) == 4) hi = P - 1 - (hi & 0xffff);
        if ((it & 12) == 8) hi = hi & 0xffff;
        if (redc(lo, hi) != redc_reference(lo, hi)) nb++;
        fe a = lo % P, b = hi;
        u128 t = mul_wide(a, b);
        if (redc(t.lo, t.hi) != redc_reference(t.lo, t.hi)) nb++;
    }
    if (nb) atomicAdd(bad, nb);
}

__global__ void field_selfcheck_kernel(unsigned long long *bad, unsigned long long seed) {
    const unsigned long long nb = field_selfcheck(seed + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ULL, 64);
    if (nb) atomicAdd(bad, nb);
}

}  // namespace

long long field_selftest(Stream &st) {
    unsigned long long *d, h = 0;
    CSG_CUDA(cudaMalloc((void **)&d, 8));
    CSG_CUDA(cudaMemsetAsync(d, 0, 8, st.s));
    CSG_LAUNCH(st, field_selfcheck_kernel, 256, 128, 0, d, 0x5eedULL);
    CSG_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, st.s));
    CSG_CUDA(cudaStreamSynchronize(st.s));
    cudaFree(d);
    return (long long)h;
}

long long redc_selftest(Stream &st) {
    unsigned long long *d, h = 0;
    CSG_CUDA(cudaMalloc((void **)&d, 8));
    CSG_CUDA(cudaMemsetAsync(d, 0, 8, st.s));
    CSG_LAUNCH(st, redc_selftest_kernel, 1024, 256, 0, d, 12345ULL);
    CSG_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, st.s));
    CSG_CUDA(cudaStreamSynchronize(st.s));
    cudaFree(d);
    return (long long)h;
}

void coset_major_to_natural(const fe *lde, unsigned width, unsigned ncosets, size_t n, uint64_t *out, Stream &st) {
    dim3 grid((unsigned)((n * ncosets + 255) / 256), width);
    CSG_LAUNCH(st, coset_to_natural_kernel, grid, 256, 0, lde, width, ncosets, (unsigned long long)n, out);
}

void to_montgomery(const uint64_t *in, fe *out, size_t count, Stream &st, bool already_montgomery) {
    CSG_LAUNCH(st, to_mont_kernel, grid_for(count, 256), 256, 0, in, out, (unsigned long long)count, already_montgomery ? ONE : R2);
}
void from_montgomery(const fe *in, uint64_t *out, size_t count, Stream &st) {
    CSG_LAUNCH(st, from_mont_kernel, grid_for(count, 256), 256, 0, in, out, (unsigned long long)count);
}

void composition_columns(const fe *e, fe *cols, size_t n, unsigned ce, const fe *mat_host, Stream &st, size_t col_stride) {
    if (ce > 16) throw std::runtime_error("constraint blowup above 16 is not supported");
    CrossMat M;
    for (unsigned i = 0; i < ce * ce; i++) M.m[i] = mat_host[i];
    CSG_LAUNCH(st, composition_columns_kernel, (unsigned)((n + 255) / 256), 256, 0, e, cols, (unsigned long long)n, ce, M,
               (unsigned long long)(col_stride ? col_stride : n));
}

void coset_mix(const fe *in, fe *out, size_t n, unsigned L, size_t npolys, const fe *mat_host, Stream &st) {
    if (L > 16) throw std::runtime_error("at most 16 cosets can be mixed");
    CrossMat M;
    for (unsigned i = 0; i < L * L; i++) M.m[i] = mat_host[i];
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)npolys);
    CSG_LAUNCH(st, coset_mix_kernel, grid, 256, 0, in, out, (unsigned long long)n, L, M);
}
void coset_mix_sharded(const fe *in, fe *out, size_t n, unsigned L, unsigned Ll, size_t rank_stride, size_t npolys, const fe *mat_host, Stream &st) {
    if (L > 16 || Ll == 0 || L % Ll) throw std::runtime_error("bad coset counts for the sharded mix");
    CrossMat M;
    for (unsigned i = 0; i < Ll * L; i++) M.m[i] = mat_host[i];
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)npolys);
    CSG_LAUNCH(st, coset_mix_sharded_kernel, grid, 256, 0, in, out, (unsigned long long)n, L, Ll, (unsigned long long)rank_stride, M);
}
void sum_slices(const fe *in, fe *out, size_t stride, unsigned count, Stream &st) {
    CSG_LAUNCH(st, sum_slices_kernel, (unsigned)((stride + 255) / 256), 256, 0, in, out, (unsigned long long)stride, count);
}

void eval_polys_at(const fe *polys, size_t stride, size_t ncols, size_t n, const fe *points_host, size_t npoints, fe *values_host,
                   DBuf<fe> &scratch, Stream &st) {
    if (npoints > 4) throw std::runtime_error("at most 4 evaluation points per call");
    const unsigned nblk = (unsigned)((n + EVAL_THREADS * EVAL_CHUNK - 1) / (EVAL_THREADS * EVAL_CHUNK));
    const size_t total = (size_t)nblk * ncols * npoints;
    scratch.reserve(total);
    EvalPoints pts{};
    for (size_t p = 0; p < npoints; p++) { pts.z[p] = points_host[p]; pts.z_stride[p] = f63::pow(points_host[p], EVAL_THREADS); }
    dim3 grid(nblk, (unsigned)ncols);
    switch (npoints) {
    case 1: CSG_LAUNCH(st, eval_polys_kernel<1>, grid, EVAL_THREADS, 0, polys, (unsigned long long)stride, (unsigned long long)n, pts, scratch.p); break;
    case 2: CSG_LAUNCH(st, eval_polys_kernel<2>, grid, EVAL_THREADS, 0, polys, (unsigned long long)stride, (unsigned long long)n, pts, scratch.p); break;
    case 3: CSG_LAUNCH(st, eval_polys_kernel<3>, grid, EVAL_THREADS, 0, polys, (unsigned long long)stride, (unsigned long long)n, pts, scratch.p); break;
    default: CSG_LAUNCH(st, eval_polys_kernel<4>, grid, EVAL_THREADS, 0, polys, (unsigned long long)stride, (unsigned long long)n, pts, scratch.p); break;
    }
    std::vector<fe> part(total);
    CSG_CUDA(cudaMemcpyAsync(part.data(), scratch.p, total * sizeof(fe), cudaMemcpyDeviceToHost, st.s));
    CSG_CUDA(cudaStreamSynchronize(st.s));
    for (size_t pc = 0; pc < npoints * ncols; pc++) {
        fe s = 0;
        for (unsigned b = 0; b < nblk; b++) s = add(s, part[pc * nblk + b]);
        values_host[pc] = s;
    }
}

void combine_polys(const fe *polys, size_t stride, size_t ncols, size_t n, const fe *coef_host, size_t ncomb, fe *out, size_t out_stride,
                   DBuf<fe> &scratch, Stream &st) {
    scratch.reserve(ncomb * ncols);
    CSG_CUDA(cudaMemcpyAsync(scratch.p, coef_host, ncomb * ncols * sizeof(fe), cudaMemcpyHostToDevice, st.s));
    const unsigned grid = (unsigned)((n + 255) / 256);
#define CSG_COMBINE(N) CSG_LAUNCH(st, combine_polys_kernel<N>, grid, 256, 0, polys, (unsigned long long)stride, (unsigned)ncols, (unsigned long long)n, scratch.p, out, (unsigned long long)out_stride)
    switch (ncomb) {
    case 1: CSG_COMBINE(1); break;
    case 2: CSG_COMBINE(2); break;
    case 3: CSG_COMBINE(3); break;
    case 4: CSG_COMBINE(4); break;
    case 6: CSG_COMBINE(6); break;
    default: throw std::runtime_error("combine_polys: 1, 2, 3, 4 or 6 combinations per call");
    }
#undef CSG_COMBINE
}

void deep_quotients(const fe *abc, const fe *W, size_t n, const DeepArgs &a, fe *deep, Stream &st) {
    const size_t total = n * a.ncosets, per_block = (size_t)DEEP_CHUNK * DEEP_THREADS;
    CSG_LAUNCH(st, deep_quotients_kernel, (unsigned)((total + per_block - 1) / per_block), DEEP_THREADS, 0, abc, W, (unsigned long long)n, a, deep);
}

void fri_fold4(const fe *evals, size_t m, const fe *W, const FoldArgs &a, fe *out, Stream &st) {
    const size_t q = m / 4;
    CSG_LAUNCH(st, fri_fold4_kernel, (unsigned)((q + 255) / 256), 256, 0, evals, (unsigned long long)q, W, a, out);
}

void gather_rows(const fe *data, unsigned width, unsigned ncosets, size_t coset_stride, size_t col_stride, const uint32_t *pos_dev,
                 size_t npos, uint64_t *rows_dev, Stream &st, unsigned sub, size_t sub_stride) {
    if (!npos) return;
    const size_t total = npos * width;
    CSG_LAUNCH(st, gather_rows_kernel, (unsigned)((total + 255) / 256), 256, 0, data, width, ncosets, (unsigned long long)coset_stride,
               (unsigned long long)col_stride, pos_dev, (unsigned)npos, rows_dev, sub ? sub : 1u, (unsigned long long)sub_stride);
}

}  // namespace csg
