// Kernels of the combined-mode constraint evaluation, shared by constraints.cu (base field) and constraints_ext.cu
// (FieldExtension::Quadratic / Cubic: the same kernels with DEG accumulators per thread).  See constraints.cuh / airs.cuh.
#pragma once
#include <cstdlib>
#include <vector>

#include "airs.cuh"
#include "constraints.cuh"
#include "stages.cuh"

namespace csg {
using namespace f63;

namespace {
constexpr int CONS_THREADS = 128;

struct RowCtx {
    airs::Frame f;
    airs::Periodic pv;
    fe x;
};
// The row kernels read each LDE column once, a value at a time, as the straight-line constraint code reaches it: every first touch
// of a column is an exposed trip to HBM (ncu: long_scoreboard 3.4-7.6 warps per issue in the linear-rest, merge and final kernels).
// One TMA bulk prefetch per column at the top of the CTA (cp.async.bulk.prefetch.L2: no destination, no barrier, one instruction
// per 1 KB segment, issued by the first warps) brings the CTA's 128-row segment of every column it will read into L2 while the
// prologue runs; the loads then cost an L2 hit.  Segments are 1 KB-aligned (row index a multiple of 128) and clamped to the column.
__device__ __forceinline__ void prefetch_tile_l2(const fe *base, unsigned long long col_stride, unsigned col0, unsigned ncols, unsigned long long i0,
                                                 unsigned long long n) {
#ifndef CSG_NO_TILE_PREFETCH
    const unsigned long long rows = n - i0 < (unsigned long long)CONS_THREADS + 2 ? n - i0 : (unsigned long long)CONS_THREADS + 2;
    const unsigned bytes = (unsigned)(rows * sizeof(fe)) & ~15u;
    if (!bytes) return;
    for (unsigned c = threadIdx.x; c < ncols; c += CONS_THREADS)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + (unsigned long long)(col0 + c) * col_stride + i0), "r"(bytes) : "memory");
#endif
}

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
#ifndef CSG_LOW_MINBLOCKS
#define CSG_LOW_MINBLOCKS 4
#endif
#ifndef CSG_REST_MINBLOCKS
#define CSG_REST_MINBLOCKS 6
#endif
// KIND 0: Rescue residual number blockIdx.z; KIND 1: scalar-multiplication bank blockIdx.z; KIND 2: final point addition
template <int AIR, int KIND, int DEG = 1>
__global__ void __launch_bounds__(CONS_THREADS, KIND == 0 ? (DEG == 1 && AIR != airs::TRANSACTION ? CSG_RESCUE_MINBLOCKS : 4) : CSG_ECC_MINBLOCKS)
cons_item_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab,
                 fe *__restrict__ part, unsigned part_items) {
    __shared__ fe xp_s[airs::MAX_GROUPS][CONS_THREADS];   // x^adj of each degree group, one column per thread
    const unsigned kc = blockIdx.y, item = KIND == 2 ? 2 : blockIdx.z;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    RowCtx r = row_setup(A, lde, W, ptab, kc, i, xp_s);
    airs::CombT<false, DEG> C{A->alpha, A->beta, A->group, &xp_s[0][threadIdx.x], (size_t)CONS_THREADS, acc192(), nullptr, 0,
                              &A->alpha_x[0][0], &A->beta_x[0][0], (size_t)CONS_MAX_CONSTRAINTS};
    C.rt = &A->rt;
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
    mark(5);   // no separate formula kernel on this path
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

// ================================================================================================ low-degree split
// Rescue residuals, the linear rest and the curve formulas of the banks have degree < (ce/2) * n: their alpha part and
// per-group beta parts are evaluated on the EVEN ce cosets only, in split mode, interpolated there and evaluated on the odd
// cosets.  With E-valued coefficients (DEG = 2, 3) every polynomial exists once per component: index (comp * NP + p).
// The per-thread accumulators of the beta parts live in dynamic shared memory: 3 words x MAX_SPLIT_GROUPS x DEG per thread.
constexpr size_t split_smem_bytes(int deg) { return (size_t)3 * airs::MAX_SPLIT_GROUPS * deg * CONS_THREADS * sizeof(uint64_t); }

// KIND 0: Rescue residual number blockIdx.z; KIND 3: the linear rest.
// low[(((item * DEG + comp) * NP + p) * L + j) * n + i], p = 0 the alpha part, p = 1 + g the beta part of degree group g, j = kc / 2.
template <int AIR, int KIND, int DEG>
__global__ void __launch_bounds__(CONS_THREADS, CSG_LOW_MINBLOCKS)   // 5 or 6 CTAs (spills) measured no faster for either kind
cons_low_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab,
                fe *__restrict__ low, unsigned item0) {
    extern __shared__ uint64_t part_dyn[];
    const unsigned j = blockIdx.y, kc = 2 * j, item = item0 + blockIdx.z, L = A->ncosets / 2, NP = 1 + A->ngroups;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    const unsigned long long inext = (i + 1) & (n - 1);
    const fe *base = lde + A->lde_coset_stride[kc];
    if (KIND == 3) prefetch_tile_l2(base, A->col_stride, 0, A->width, blockIdx.x * (unsigned long long)CONS_THREADS, n);
    airs::Frame f{base + i, base + inext, (size_t)A->col_stride};
    airs::Periodic pv{ptab + kc * A->ptab_coset_stride, A->poff, A->pmask, (uint32_t)i};
    for (unsigned k = 0; k < 3 * airs::MAX_SPLIT_GROUPS * DEG; k++) part_dyn[k * CONS_THREADS + threadIdx.x] = 0;
    airs::CombT<true, DEG> C{A->alpha, A->beta, A->group, nullptr, 0, acc192(), &part_dyn[threadIdx.x], (size_t)CONS_THREADS,
                             &A->alpha_x[0][0], &A->beta_x[0][0], (size_t)CONS_MAX_CONSTRAINTS};
    C.rt = &A->rt;
    if (KIND == 0) airs::eval_rescue_item<AIR>((int)blockIdx.z, f, pv, C);
    else airs::eval_rest<AIR>(f, pv, C);
#pragma unroll
    for (int comp = 0; comp < DEG; comp++) {
        fe *dst = low + ((((unsigned long long)item * DEG + comp) * NP) * L + j) * n + i;
        dst[0] = comp == 0 ? C.sum.reduce() : C.sum_x[comp > 0 ? comp - 1 : 0].reduce();
        for (unsigned g = 0; g + 1 < NP; g++) dst[(unsigned long long)(1 + g) * L * n] = C.part_value((int)g, comp);
    }
}

// ---- the scalar-multiplication banks (airs.cuh, scalar_mult_bank_outputs): the merged outputs of the doubling and the
// mixed-addition formula of each bank -- 4 variants -- on the even cosets, as an alpha polynomial and one beta polynomial per
// degree group the bank's slots fall into, per component.
constexpr unsigned ECC_SPLIT_MAX_POLYS = 16;
constexpr unsigned ECC_SUPER = 512;   // row tiles per turn of a variant (cons_ecc_low_kernel)
struct EccSplitMap {
    unsigned npolys[2];                                // bank b: 1 + number of groups among its point slots
    unsigned char groups[2][airs::MAX_SPLIT_GROUPS];   // those groups
    unsigned base[4];                                  // first polynomial (per component) of variant v = 2*bank + formula
    unsigned total;
};
// eccl[((base[v] * DEG + comp * npolys[bank] + q) * L + j) * n + i]
template <int AIR, int DEG>
__global__ void __launch_bounds__(CONS_THREADS, CSG_ECC_MINBLOCKS)
cons_ecc_low_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ ptab, EccSplitMap M, fe *__restrict__ eccl) {
    extern __shared__ uint64_t part_dyn[];
    // The two formulas of a bank read the same 18 point columns.  Launched as separate grid planes the second read came from
    // DRAM again (ncu, round 1: 3.56 GB of traffic against 1.95 GB algorithmic); with the four variants of one tile as
    // neighbouring CTAs it is an L2 hit but four instruction streams share every SM (+2 % time).  So the grid runs the variants
    // in turns over super-tiles of ECC_SUPER row tiles: one code stream at a time, and the re-read is ~16 MB later, inside L2.
    const unsigned j = blockIdx.y, kc = 2 * j, L = A->ncosets / 2;
    const unsigned rem = blockIdx.x % (4 * ECC_SUPER), v = rem / ECC_SUPER, bank = v >> 1, npb = M.npolys[bank];
    const unsigned long long n = 1ULL << A->logn, tile = (unsigned long long)(blockIdx.x / (4 * ECC_SUPER)) * ECC_SUPER + rem % ECC_SUPER;
    const unsigned long long i = tile * CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    const unsigned long long inext = (i + 1) & (n - 1);
    const fe *base = lde + A->lde_coset_stride[kc];
    airs::Frame f{base + i, base + inext, (size_t)A->col_stride};
    airs::Periodic pv{ptab + kc * A->ptab_coset_stride, A->poff, A->pmask, (uint32_t)i};
    for (unsigned k = 0; k < 3 * airs::MAX_SPLIT_GROUPS * DEG; k++) part_dyn[k * CONS_THREADS + threadIdx.x] = 0;
    airs::CombT<true, DEG> C{A->alpha, A->beta, A->group, nullptr, 0, acc192(), &part_dyn[threadIdx.x], (size_t)CONS_THREADS,
                             &A->alpha_x[0][0], &A->beta_x[0][0], (size_t)CONS_MAX_CONSTRAINTS};
    airs::eval_ecc_bank_outputs<AIR>((int)bank, (int)(v & 1), f, pv, C);
#pragma unroll
    for (int comp = 0; comp < DEG; comp++) {
        fe *dst = eccl + (((unsigned long long)M.base[v] * DEG + comp * npb) * L + j) * n + i;
        dst[0] = comp == 0 ? C.sum.reduce() : C.sum_x[comp > 0 ? comp - 1 : 0].reduce();
        for (unsigned q = 0; q + 1 < npb; q++) dst[(unsigned long long)(1 + q) * L * n] = C.part_value((int)M.groups[bank][q], comp);
    }
}
// the banks' contribution to T(x) on every ce coset from the extended formula values; hi[((comp * nhi + bank) * ncosets + kc) * n + i]
template <int AIR, int DEG>
__global__ void __launch_bounds__(CONS_THREADS)
cons_ecc_merge_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab, EccSplitMap M,
                      const fe *__restrict__ even, const fe *__restrict__ odd, fe *__restrict__ hi, unsigned nhi) {
    __shared__ fe xp_s[airs::MAX_GROUPS][CONS_THREADS];
    const unsigned kc = blockIdx.y, bank = blockIdx.z, L = A->ncosets / 2, npb = M.npolys[bank];
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    {   // this CTA's segment of the bank's point and bit columns, and of the extended formula polynomials
        const unsigned long long i0 = blockIdx.x * (unsigned long long)CONS_THREADS;
        prefetch_tile_l2(lde + A->lde_coset_stride[kc], A->col_stride, bank * (airs::PPW + 1), airs::PPW + 1, i0, n);
        const fe *lp = ((kc & 1) ? odd : even) + (unsigned long long)(kc >> 1) * n;
        for (unsigned formula = 0; formula < 2; formula++)
            prefetch_tile_l2(lp + (unsigned long long)M.base[2 * bank + formula] * DEG * L * n, (unsigned long long)L * n, 0, DEG * npb, i0, n);
    }
    RowCtx r = row_setup(A, lde, W, ptab, kc, i, xp_s);
    airs::CombT<false, DEG> C{A->alpha, A->beta, A->group, &xp_s[0][threadIdx.x], (size_t)CONS_THREADS, acc192(), nullptr, 0,
                              &A->alpha_x[0][0], &A->beta_x[0][0], (size_t)CONS_MAX_CONSTRAINTS};
    const fe *low = ((kc & 1) ? odd : even) + (unsigned long long)(kc >> 1) * n + i;
#pragma unroll
    for (int comp = 0; comp < DEG; comp++) {
        fe merged[2];
        for (unsigned formula = 0; formula < 2; formula++) {
            const fe *p = low + ((unsigned long long)M.base[2 * bank + formula] * DEG + comp * npb) * L * n;
            acc192 s;
            for (unsigned q = 0; q + 1 < npb; q++) s.mac(xp_s[M.groups[bank][q]][threadIdx.x], p[(unsigned long long)(1 + q) * L * n]);
            merged[formula] = add(s.reduce(), p[0]);
        }
        hi[(((unsigned long long)comp * nhi + bank) * A->ncosets + kc) * n + i] = airs::eval_ecc_bank_merge<AIR>((int)bank, r.f, r.pv, C, comp, merged[0], merged[1]);
    }
}

// T(x) from its pieces on every ce coset, then divisors and boundary constraints, per component.
//   low_even / low_odd: [DEG][NP][L][n] values of A and B_g on the even / odd cosets; hi: [DEG][nhi][ncosets][n] partial sums
//   of the curve items, already complete.
template <int DEG>
__global__ void __launch_bounds__(CONS_THREADS)
cons_final_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ apoly,
                  const fe *__restrict__ low_even, const fe *__restrict__ low_odd, const fe *__restrict__ hi, unsigned nhi,
                  const fe *__restrict__ binv, fe *__restrict__ out) {
    const unsigned kc = blockIdx.y, L = A->ncosets / 2, NP = 1 + A->ngroups;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    {   // the CTA's segment of every array it sums: A and the B_g of every component, the curve partial sums, the divisor inverses
        const unsigned long long i0 = blockIdx.x * (unsigned long long)CONS_THREADS;
        prefetch_tile_l2(((kc & 1) ? low_odd : low_even) + (unsigned long long)(kc >> 1) * n, (unsigned long long)L * n, 0, DEG * NP, i0, n);
        prefetch_tile_l2(hi + (unsigned long long)kc * n, (unsigned long long)A->ncosets * n, 0, DEG * nhi, i0, n);
        prefetch_tile_l2(binv + (unsigned long long)kc * n, (unsigned long long)A->ncosets * n, 0, A->nbgroups, i0, n);
    }
    const fe x = mul(A->shift[kc], W[i]);
    const fe *low0 = ((kc & 1) ? low_odd : low_even) + (unsigned long long)(kc >> 1) * n + i;   // component c: + c * NP * L * n
    const unsigned long long comp_stride = (unsigned long long)NP * L * n;
    acc192 s[DEG];
    for (unsigned g = 0; g < A->ngroups; g++) {
        const fe xp = mul(A->shift_adj[kc][g], W[(A->adj_mod[g] * i) & (n - 1)]);
#pragma unroll
        for (int comp = 0; comp < DEG; comp++) s[comp].mac(xp, low0[comp * comp_stride + (unsigned long long)(1 + g) * L * n]);
    }
    fe res[DEG];
#pragma unroll
    for (int comp = 0; comp < DEG; comp++) {
        fe t = add(s[comp].reduce(), low0[comp * comp_stride]);
        for (unsigned p = 0; p < nhi; p++) t = add(t, hi[(((unsigned long long)comp * nhi + p) * A->ncosets + kc) * n + i]);
        res[comp] = mul(mul(t, sub(x, A->g_last)), A->zinv[kc]);
    }
    const fe *cur = lde + A->lde_coset_stride[kc] + i;
    unsigned a = 0;
    for (unsigned g = 0; g < A->nbgroups; g++) {
        const fe xpb = mul(A->b_shift_adj[kc][g], W[(A->b_adj_mod[g] * i) & (n - 1)]);
        acc192 b[DEG];
        for (; a < A->nassertions && A->a_group[a] == g; a++) {
            fe v = A->a_value[a];
            if (A->a_poly_len[a] > 1) {
                const fe *poly = apoly + A->a_poly_off[a];
                const fe y = mul(x, A->a_xoff[a]);
                v = 0;
                for (unsigned m = A->a_poly_len[a]; m-- > 0;) v = add(mul(v, y), poly[m]);
            }
            const fe dv = sub(cur[(unsigned long long)A->a_col[a] * A->col_stride], v);
            b[0].mac(add(A->a_alpha[a], mul(A->a_beta[a], xpb)), dv);
#pragma unroll
            for (int comp = 1; comp < DEG; comp++) b[comp].mac(add(A->a_alpha_x[comp - 1][a], mul(A->a_beta_x[comp - 1][a], xpb)), dv);
        }
        const fe bi = binv[((unsigned long long)g * A->ncosets + kc) * n + i];
#pragma unroll
        for (int comp = 0; comp < DEG; comp++) res[comp] = add(res[comp], mul(b[comp].reduce(), bi));
    }
#pragma unroll
    for (int comp = 0; comp < DEG; comp++) out[((unsigned long long)comp * A->ncosets + kc) * n + i] = res[comp];
}

// the split needs the 8 ce cosets of the transaction / Schnorr AIR; a rank of a sharded proof must own whole even/odd pairs
inline bool split_applies(const ConsArgs &h, const SplitExchange *xch) {
    const unsigned world = xch ? xch->world : 1;
    return h.ncosets * world == 8 && h.ncosets % 2 == 0 && h.ngroups <= (unsigned)airs::MAX_SPLIT_GROUPS;
}

template <int AIR, int DEG = 1>
void launch_split(const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly, fe *part, fe *out, Stream &st,
                  cudaEvent_t *ev, const RootTable &rt, NttScratch &sc, const SplitExchange *xch = nullptr) {
    const unsigned long long n = 1ULL << h.logn;
    // ce, L: the cosets of THIS rank (all of them without an exchange); Lg: even cosets of the whole proof
    const unsigned gx = (unsigned)((n + CONS_THREADS - 1) / CONS_THREADS), ce = h.ncosets, L = ce / 2, NP = 1 + h.ngroups, NPD = DEG * NP;
    const unsigned G = xch ? xch->world : 1, Lg = L * G, rank = xch ? xch->rank : 0;
    constexpr int NR = airs::Items<AIR>::rescue, NE = airs::Items<AIR>::ecc;
    const size_t slab = (size_t)NPD * L * n;
    fe *low_parts = part, *low_sum = low_parts + (size_t)(NR + 1) * slab, *low_coef = low_sum + slab, *low_mix = low_coef + slab,
       *low_odd = low_mix + slab, *hi = low_odd + slab, *binv = hi + (size_t)DEG * NE * ce * n;
    // interpolation on the even cosets, L x L mix of the per-coset coefficient sets, evaluation on the odd cosets, for `np`
    // polynomials laid out [p][L][n]
    const fe w2l = root_of_unity(ilog2(2 * Lg)), linv = inv(to_mont(Lg));
    std::vector<fe> mix(L * Lg);   // the rows of the Lg x Lg map that produce this rank's odd cosets
    for (unsigned jl = 0; jl < L; jl++)
        for (unsigned je = 0; je < Lg; je++) {
            const unsigned jo = rank * L + jl, e = (2 * (jo + Lg - je) + 1) % (2 * Lg);   // 2 (j' - j) + 1 mod 2L
            fe acc = 0, step = f63::pow(w2l, e), cur = ONE;
            for (unsigned t = 0; t < Lg; t++) { acc = add(acc, cur); cur = mul(cur, step); }
            mix[jl * Lg + je] = mul(acc, linv);
        }
    std::vector<fe> even_inv(L);
    for (unsigned j = 0; j < L; j++) even_inv[j] = inv(h.shift[2 * j]);
    auto extend = [&](const fe *even, fe *coef, fe *mixed, fe *odd, unsigned np) {
        // (the shift vectors are pageable host memory: their copies are staged before cudaMemcpyAsync returns, so nothing
        // here waits for the stream -- two synchronisations per proof less, 0.2 ms of a one-transaction proof with the
        // inversions that used to be repeated per polynomial)
        std::vector<fe> sinv(np * L), sodd(np * L);
        for (unsigned p = 0; p < np; p++)
            for (unsigned j = 0; j < L; j++) { sinv[p * L + j] = even_inv[j]; sodd[p * L + j] = h.shift[2 * j + 1]; }
        if (xch) {   // own even cosets into this rank's slice, all-gather, then only the rows of the mix this rank needs
            const size_t slice = (size_t)np * L * n;
            if (slice * G > xch->buf_elems) throw std::runtime_error("exchange buffer of the sharded split is too small");
            coset_intt_columns(rt, sc, even, n, xch->buf + rank * slice, n, h.logn, sinv.data(), (size_t)np * L, st);
            xch->gather(xch->self, xch->buf, slice * sizeof(fe));
            coset_mix_sharded(xch->buf, mixed, n, Lg, L, slice, np, mix.data(), st);
        } else {
            coset_intt_columns(rt, sc, even, n, coef, n, h.logn, sinv.data(), (size_t)np * L, st);
            coset_mix(coef, mixed, n, L, np, mix.data(), st);
        }
        coset_ntt_entries(rt, sc, mixed, odd, (size_t)np * L, h.logn, sodd.data(), st);
    };
    // the banks' formula outputs in split mode: which degree groups the point slots of each bank fall into
    static const bool ecc_split = getenv("CSG_NO_ECC_SPLIT") == nullptr;
    EccSplitMap M{};
    for (unsigned b = 0; b < 2; b++) {
        unsigned cnt = 0;
        for (unsigned sl = b * (airs::PPW + 1); sl < b * (airs::PPW + 1) + airs::PPW; sl++) {
            bool seen = false;
            for (unsigned q = 0; q < cnt; q++) seen = seen || M.groups[b][q] == h.group[sl];
            if (!seen) M.groups[b][cnt++] = h.group[sl];
        }
        M.npolys[b] = 1 + cnt;
    }
    for (unsigned v = 0; v < 4; v++) { M.base[v] = M.total; M.total += M.npolys[v >> 1]; }
    if (M.total > ECC_SPLIT_MAX_POLYS) throw std::runtime_error("too many degree groups for the curve split");
    const size_t eslab = (size_t)DEG * M.total * L * n;
    fe *eccl_even = binv + (size_t)CONS_MAX_BGROUPS * ce * n, *eccl_coef = eccl_even + eslab, *eccl_mix = eccl_coef + eslab, *eccl_odd = eccl_mix + eslab;
    const size_t smem = split_smem_bytes(DEG);
    if (smem > 48 * 1024) {
        CSG_CUDA(cudaFuncSetAttribute((cons_low_kernel<AIR, 0, DEG>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CSG_CUDA(cudaFuncSetAttribute((cons_low_kernel<AIR, 3, DEG>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CSG_CUDA(cudaFuncSetAttribute((cons_ecc_low_kernel<AIR, DEG>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    auto mark = [&](int k) { if (ev) CSG_CUDA(cudaEventRecord(ev[k], st.s)); };
    host_mark("  cons: split setup");
    mark(0);
    CSG_LAUNCH(st, (cons_low_kernel<AIR, 0, DEG>), dim3(gx, L, NR), CONS_THREADS, smem, args_dev, lde, W, ptab, low_parts, 0u);
    mark(1);
    if (ecc_split) {
        CSG_LAUNCH(st, (cons_ecc_low_kernel<AIR, DEG>), dim3((gx + ECC_SUPER - 1) / ECC_SUPER * 4 * ECC_SUPER, L, 1), CONS_THREADS, smem, args_dev, lde, ptab, M, eccl_even);
        mark(5);   // ev[1] .. ev[5]: the curve-formula kernel alone, the largest single launch of a proof
        host_mark("  cons: low + curve-low launched");
        extend(eccl_even, eccl_coef, eccl_mix, eccl_odd, DEG * M.total);
        host_mark("  cons: curve polynomials extended");
        CSG_LAUNCH(st, (cons_ecc_merge_kernel<AIR, DEG>), dim3(gx, ce, 2), CONS_THREADS, 0, args_dev, lde, W, ptab, M, (const fe *)eccl_even, (const fe *)eccl_odd, hi,
                   (unsigned)NE);
    } else {
        mark(5);
        CSG_LAUNCH(st, (cons_item_kernel<AIR, 1, DEG>), dim3(gx, ce, 2), CONS_THREADS, 0, args_dev, lde, W, ptab, hi, (unsigned)NE);
    }
    mark(2);
    CSG_LAUNCH(st, (cons_item_kernel<AIR, 2, DEG>), dim3(gx, ce, 1), CONS_THREADS, 0, args_dev, lde, W, ptab, hi, (unsigned)NE);
    mark(3);
    CSG_LAUNCH(st, (cons_low_kernel<AIR, 3, DEG>), dim3(gx, L, 1), CONS_THREADS, smem, args_dev, lde, W, ptab, low_parts, (unsigned)NR);
    sum_slices(low_parts, low_sum, slab, NR + 1, st);
    host_mark("  cons: merge, final item, rest launched");
    extend(low_sum, low_coef, low_mix, low_odd, NPD);   // interpolate on the even cosets, evaluate on the odd ones
    host_mark("  cons: low polynomials extended");
    if (h.nbgroups > 0)
        CSG_LAUNCH(st, boundary_inverse_kernel, dim3((unsigned)((n + INV_CHUNK * INV_THREADS - 1) / (INV_CHUNK * INV_THREADS)), ce, h.nbgroups),
                   INV_THREADS, 0, args_dev, W, binv);
    CSG_LAUNCH(st, cons_final_kernel<DEG>, dim3(gx, ce), CONS_THREADS, 0, args_dev, lde, W, apoly, (const fe *)low_sum, (const fe *)low_odd, (const fe *)hi,
               (unsigned)NE, (const fe *)binv, out);
    mark(4);
}

}  // namespace
}  // namespace csg
