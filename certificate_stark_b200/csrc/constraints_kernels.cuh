// Kernels of the combined-mode constraint evaluation, shared by constraints.cu (base field) and constraints_ext.cu
// (FieldExtension::Quadratic / Cubic: the same kernels with DEG accumulators per thread).  See constraints.cuh / airs.cuh.
#pragma once
#include "airs.cuh"
#include "constraints.cuh"

namespace csg {
using namespace f63;

namespace {
constexpr int CONS_THREADS = 128;

struct RowCtx {
    airs::Frame f;
    airs::Periodic pv;
    fe x;
};
// common prologue: frame, periodic accessor, x and the x^adj table of this thread's row
__device__ __forceinline__ RowCtx row_setup(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W,
                                            const fe *__restrict__ ptab, unsigned kc, unsigned long long i, fe (*xp_s)[CONS_THREADS]) {
    const unsigned long long n = 1ULL << A->logn, inext = (i + 1) & (n - 1);
    const fe *base = lde + A->lde_coset_stride[kc];
    RowCtx r{airs::Frame{base + i, base + inext, (size_t)A->col_stride},
             airs::Periodic{ptab + kc * A->ptab_coset_stride, A->poff, A->pmask, (uint32_t)i}, mul(A->shift[kc], W[i])};
    for (unsigned g = 0; g < A->ngroups; g++) xp_s[g][threadIdx.x] = mul(A->shift_adj[kc][g], W[(A->adj_mod[g] * i) & (n - 1)]);
    return r;
}

// Inverse evaluations of the boundary divisors, 1 / (x^steps_g - offset_g) for every row of every ce coset, by batch
// inversion (Montgomery's trick): a thread inverts INV_CHUNK values with one field inversion and 3 multiplications each,
// instead of one ~90-multiplication inversion per row and divisor in the row kernel.
constexpr int INV_CHUNK = 16, INV_THREADS = 128;
__global__ void __launch_bounds__(INV_THREADS) boundary_inverse_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ W, fe *__restrict__ binv) {
    const unsigned kc = blockIdx.y, g = blockIdx.z;
    const unsigned long long n = 1ULL << A->logn;
    // element c of this thread: row i = block_base + c * INV_THREADS + tid (coalesced across the warp)
    const unsigned long long base = blockIdx.x * (unsigned long long)(INV_CHUNK * INV_THREADS) + threadIdx.x;
    const fe shift = A->b_steps[g] == 1 ? A->shift[kc] : A->b_shift_steps[kc][g], off = A->b_offset[g];
    fe d[INV_CHUNK], pre[INV_CHUNK];
    fe acc = ONE;
#pragma unroll
    for (int c = 0; c < INV_CHUNK; c++) {
        const unsigned long long i = base + (unsigned long long)c * INV_THREADS;
        d[c] = i < n ? sub(mul(shift, W[(A->b_steps[g] * i) & (n - 1)]), off) : ONE;
        pre[c] = acc;
        acc = mul(acc, d[c]);
    }
    fe ainv = inv(acc);
    fe *out = binv + ((unsigned long long)g * A->ncosets + kc) * n;
#pragma unroll
    for (int c = INV_CHUNK - 1; c >= 0; c--) {
        const unsigned long long i = base + (unsigned long long)c * INV_THREADS;
        if (i < n) out[i] = mul(ainv, pre[c]);
        ainv = mul(ainv, d[c]);
    }
}

#ifndef CSG_ECC_MINBLOCKS
#define CSG_ECC_MINBLOCKS 3
#endif
#ifndef CSG_RESCUE_MINBLOCKS
#define CSG_RESCUE_MINBLOCKS 6
#endif
#ifndef CSG_REST_MINBLOCKS
#define CSG_REST_MINBLOCKS 6
#endif
// KIND 0: Rescue residual number blockIdx.z; KIND 1: scalar-multiplication bank blockIdx.z; KIND 2: final point addition
template <int AIR, int KIND, int DEG = 1>
__global__ void __launch_bounds__(CONS_THREADS, KIND == 0 ? (DEG == 1 ? CSG_RESCUE_MINBLOCKS : 4) : CSG_ECC_MINBLOCKS)
cons_item_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab,
                 fe *__restrict__ part, unsigned part_items) {
    __shared__ fe xp_s[airs::MAX_GROUPS][CONS_THREADS];   // x^adj of each degree group, one column per thread
    const unsigned kc = blockIdx.y, item = KIND == 2 ? 2 : blockIdx.z;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    RowCtx r = row_setup(A, lde, W, ptab, kc, i, xp_s);
    airs::CombT<false, DEG> C{A->alpha, A->beta, A->group, &xp_s[0][threadIdx.x], (size_t)CONS_THREADS, acc192(), nullptr, 0,
                              &A->alpha_x[0][0], &A->beta_x[0][0], (size_t)CONS_MAX_CONSTRAINTS};
    if (KIND == 0) airs::eval_rescue_item<AIR>((int)item, r.f, r.pv, C);
    else if (KIND == 1) airs::eval_ecc_bank<AIR>((int)item, r.f, r.pv, C);
    else airs::eval_ecc_final<AIR>(r.f, r.pv, C);
    // component j of item `item`: part[((j * gridDim-items + item) * ncosets + kc) * n + i]; nitems travels in part_items
    part[((unsigned long long)item * A->ncosets + kc) * n + i] = C.sum.reduce();
#pragma unroll
    for (int j = 1; j < DEG; j++) part[(((unsigned long long)j * part_items + item) * A->ncosets + kc) * n + i] = C.sum_x[j - 1].reduce();
}

template <int AIR, int DEG = 1>
__global__ void __launch_bounds__(CONS_THREADS, DEG == 1 ? CSG_REST_MINBLOCKS : 4)
cons_rest_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab,
                 const fe *__restrict__ apoly, const fe *__restrict__ part, unsigned nparts, const fe *__restrict__ binv, fe *__restrict__ out) {
    __shared__ fe xp_s[airs::MAX_GROUPS][CONS_THREADS];
    const unsigned kc = blockIdx.y;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    RowCtx r = row_setup(A, lde, W, ptab, kc, i, xp_s);
    airs::CombT<false, DEG> C{A->alpha, A->beta, A->group, &xp_s[0][threadIdx.x], (size_t)CONS_THREADS, acc192(), nullptr, 0,
                              &A->alpha_x[0][0], &A->beta_x[0][0], (size_t)CONS_MAX_CONSTRAINTS};
    airs::eval_rest<AIR>(r.f, r.pv, C);
    const fe x = r.x;
    fe res[DEG];
#pragma unroll
    for (int j = 0; j < DEG; j++) {
        fe t = j == 0 ? C.sum.reduce() : C.sum_x[j > 0 ? j - 1 : 0].reduce();
        for (unsigned p = 0; p < nparts; p++) t = add(t, part[(((unsigned long long)j * nparts + p) * A->ncosets + kc) * n + i]);
        res[j] = mul(mul(t, sub(x, A->g_last)), A->zinv[kc]);
    }
    unsigned a = 0;
    for (unsigned g = 0; g < A->nbgroups; g++) {
        const fe xpb = mul(A->b_shift_adj[kc][g], W[(A->b_adj_mod[g] * i) & (n - 1)]);
        acc192 s[DEG];
        for (; a < A->nassertions && A->a_group[a] == g; a++) {
            fe v = A->a_value[a];
            if (A->a_poly_len[a] > 1) {   // Assertion::sequence: value polynomial evaluated at x * g^-first_step
                const fe *poly = apoly + A->a_poly_off[a];
                const fe y = mul(x, A->a_xoff[a]);
                v = 0;
                for (unsigned m = A->a_poly_len[a]; m-- > 0;) v = add(mul(v, y), poly[m]);
            }
            const fe dv = sub(r.f.cur(A->a_col[a]), v);
            s[0].mac(add(A->a_alpha[a], mul(A->a_beta[a], xpb)), dv);
#pragma unroll
            for (int j = 1; j < DEG; j++) s[j].mac(add(A->a_alpha_x[j - 1][a], mul(A->a_beta_x[j - 1][a], xpb)), dv);
        }
        const fe bi = binv[((unsigned long long)g * A->ncosets + kc) * n + i];
#pragma unroll
        for (int j = 0; j < DEG; j++) res[j] = add(res[j], mul(s[j].reduce(), bi));
    }
#pragma unroll
    for (int j = 0; j < DEG; j++) out[((unsigned long long)j * A->ncosets + kc) * n + i] = res[j];
}

template <int AIR, int DEG = 1>
void launch(const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly, fe *part, fe *out, Stream &st,
            cudaEvent_t *ev) {
    // scratch layout: [items partial sums][boundary-divisor inverses], each ncosets * n elements per entry
    const unsigned long long n = 1ULL << h.logn;
    const unsigned gx = (unsigned)((n + CONS_THREADS - 1) / CONS_THREADS);
    constexpr int NR = airs::Items<AIR>::rescue, NE = airs::Items<AIR>::ecc;
    auto mark = [&](int k) { if (ev) CSG_CUDA(cudaEventRecord(ev[k], st.s)); };
    mark(0);
    if (NR > 0) CSG_LAUNCH(st, (cons_item_kernel<AIR, 0, DEG>), dim3(gx, h.ncosets, NR > 0 ? NR : 1), CONS_THREADS, 0, args_dev, lde, W, ptab, part, (unsigned)(NR + NE));
    mark(1);
    fe *ecc_part = part + (size_t)NR * h.ncosets * n;
    // two scalar-multiplication banks, then the final addition (its own kernel: different code, fewer registers)
    if (NE > 0) CSG_LAUNCH(st, (cons_item_kernel<AIR, 1, DEG>), dim3(gx, h.ncosets, 2), CONS_THREADS, 0, args_dev, lde, W, ptab, ecc_part, (unsigned)(NR + NE));
    mark(2);
    if (NE > 0) CSG_LAUNCH(st, (cons_item_kernel<AIR, 2, DEG>), dim3(gx, h.ncosets, 1), CONS_THREADS, 0, args_dev, lde, W, ptab, ecc_part, (unsigned)(NR + NE));
    mark(3);
    fe *binv = part + (size_t)DEG * (NR + NE) * h.ncosets * n;
    if (h.nbgroups > 0)
        CSG_LAUNCH(st, boundary_inverse_kernel, dim3((unsigned)((n + INV_CHUNK * INV_THREADS - 1) / (INV_CHUNK * INV_THREADS)), h.ncosets, h.nbgroups),
                   INV_THREADS, 0, args_dev, W, binv);
    CSG_LAUNCH(st, (cons_rest_kernel<AIR, DEG>), dim3(gx, h.ncosets), CONS_THREADS, 0, args_dev, lde, W, ptab, apoly, (const fe *)part, (unsigned)(NR + NE),
               (const fe *)binv, out);
    mark(4);
}

}  // namespace
}  // namespace csg
