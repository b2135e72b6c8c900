// See constraints.cuh / airs.cuh.  Columns are read coalesced along the row index.
//
// Launch structure per proof: one `cons_item_kernel` over (rows, ce cosets, heavy items) -- a thread evaluates ONE Rescue
// residual or ONE curve item of one row and writes its partial sum of T(x) -- then `cons_rest_kernel` over (rows, cosets)
// adds the cheap linear constraints, sums the partials, applies the divisors and the boundary constraints and writes the
// merged column.  Splitting the row's work this way keeps each kernel's code and register footprint small (the
// monolithic version ran at 8 warps/SM with 20 % of its stalls on instruction fetch) and multiplies the parallelism.
#include <cstdlib>
#include <vector>

#include "airs.cuh"
#include "constraints.cuh"
#include "constraints_kernels.cuh"
#include "stages.cuh"

namespace csg {
using namespace f63;

namespace {
// ---- low-degree split (constraints.cuh): Rescue residuals and the linear rest on the EVEN ce cosets only, in split mode.
// KIND 0: Rescue residual number blockIdx.z; KIND 3: the linear rest.  Writes 1 + ngroups partial sums per row:
// low[((item * NP + p) * L + j) * n + i], p = 0 the alpha part, p = 1 + g the beta part of degree group g, j = kc / 2.
template <int AIR, int KIND>
__global__ void __launch_bounds__(CONS_THREADS, KIND == 0 ? 4 : 4)
cons_low_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab,
                fe *__restrict__ low, unsigned item0) {
    const unsigned j = blockIdx.y, kc = 2 * j, item = item0 + blockIdx.z, L = A->ncosets / 2, NP = 1 + A->ngroups;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    const unsigned long long inext = (i + 1) & (n - 1);
    const fe *base = lde + A->lde_coset_stride[kc];
    airs::Frame f{base + i, base + inext, (size_t)A->col_stride};
    airs::Periodic pv{ptab + kc * A->ptab_coset_stride, A->poff, A->pmask, (uint32_t)i};
    __shared__ uint64_t part_s[3 * airs::MAX_SPLIT_GROUPS][CONS_THREADS];   // the per-group beta accumulators of this thread: one column
    for (unsigned k = 0; k < 3 * (NP - 1); k++) part_s[k][threadIdx.x] = 0;
    airs::SplitComb C{A->alpha, A->beta, A->group, nullptr, 0, acc192(), &part_s[0][threadIdx.x], (size_t)CONS_THREADS};
    if (KIND == 0) airs::eval_rescue_item<AIR>((int)blockIdx.z, f, pv, C);
    else airs::eval_rest<AIR>(f, pv, C);
    fe *dst = low + (((unsigned long long)item * NP) * L + j) * n + i;
    dst[0] = C.sum.reduce();
    for (unsigned g = 0; g + 1 < NP; g++) dst[(unsigned long long)(1 + g) * L * n] = C.part_value((int)g);
}

// ---- low-degree split of the scalar-multiplication banks (airs.cuh, scalar_mult_bank_outputs): the merged outputs of the
// doubling and the mixed-addition formula of each bank -- 4 variants -- on the even cosets, as an alpha polynomial and one
// beta polynomial per degree group the bank's slots fall into.
constexpr unsigned ECC_SPLIT_MAX_POLYS = 16;
struct EccSplitMap {
    unsigned npolys[2];                                // bank b: 1 + number of groups among its point slots
    unsigned char groups[2][airs::MAX_SPLIT_GROUPS];   // those groups
    unsigned base[4];                                  // first polynomial of variant v = 2*bank + formula
    unsigned total;
};
#ifndef CSG_ECC_LOW_MINBLOCKS
#define CSG_ECC_LOW_MINBLOCKS 3
#endif
template <int AIR>
__global__ void __launch_bounds__(CONS_THREADS, CSG_ECC_LOW_MINBLOCKS)
cons_ecc_low_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ ptab, EccSplitMap M, fe *__restrict__ eccl) {
    const unsigned j = blockIdx.y, kc = 2 * j, v = blockIdx.z, bank = v >> 1, L = A->ncosets / 2, NP = 1 + A->ngroups;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    const unsigned long long inext = (i + 1) & (n - 1);
    const fe *base = lde + A->lde_coset_stride[kc];
    airs::Frame f{base + i, base + inext, (size_t)A->col_stride};
    airs::Periodic pv{ptab + kc * A->ptab_coset_stride, A->poff, A->pmask, (uint32_t)i};
    __shared__ uint64_t part_s[3 * airs::MAX_SPLIT_GROUPS][CONS_THREADS];
    for (unsigned k = 0; k < 3 * (NP - 1); k++) part_s[k][threadIdx.x] = 0;
    airs::SplitComb C{A->alpha, A->beta, A->group, nullptr, 0, acc192(), &part_s[0][threadIdx.x], (size_t)CONS_THREADS};
    airs::eval_ecc_bank_outputs<AIR>((int)bank, (int)(v & 1), f, pv, C);
    fe *dst = eccl + ((unsigned long long)M.base[v] * L + j) * n + i;
    dst[0] = C.sum.reduce();
    for (unsigned q = 0; q + 1 < M.npolys[bank]; q++) dst[(unsigned long long)(1 + q) * L * n] = C.part_value((int)M.groups[bank][q]);
}
// the banks' contribution to T(x) on every ce coset from the extended formula values; hi[(bank * ncosets + kc) * n + i]
template <int AIR>
__global__ void __launch_bounds__(CONS_THREADS)
cons_ecc_merge_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ ptab, EccSplitMap M,
                      const fe *__restrict__ even, const fe *__restrict__ odd, fe *__restrict__ hi) {
    __shared__ fe xp_s[airs::MAX_GROUPS][CONS_THREADS];
    const unsigned kc = blockIdx.y, bank = blockIdx.z, L = A->ncosets / 2;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    RowCtx r = row_setup(A, lde, W, ptab, kc, i, xp_s);
    airs::Comb C{A->alpha, A->beta, A->group, &xp_s[0][threadIdx.x], (size_t)CONS_THREADS, acc192(), nullptr, 0};
    const fe *low = ((kc & 1) ? odd : even) + (unsigned long long)(kc >> 1) * n + i;
    fe merged[2];
    for (unsigned formula = 0; formula < 2; formula++) {
        const fe *p = low + (unsigned long long)M.base[2 * bank + formula] * L * n;
        acc192 s;
        for (unsigned q = 0; q + 1 < M.npolys[bank]; q++) s.mac(xp_s[M.groups[bank][q]][threadIdx.x], p[(unsigned long long)(1 + q) * L * n]);
        merged[formula] = add(s.reduce(), p[0]);
    }
    hi[((unsigned long long)bank * A->ncosets + kc) * n + i] = airs::eval_ecc_bank_merge<AIR>((int)bank, r.f, r.pv, C, merged[0], merged[1]);
}

// T(x) from its pieces on every ce coset, then divisors and boundary constraints.
//   low_even / low_odd: [NP][L][n] values of A and B_g on the even / odd cosets; hi: [nhi][ncosets][n] partial sums of the
//   high-degree items (curve arithmetic), already complete.
__global__ void __launch_bounds__(CONS_THREADS)
cons_final_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W, const fe *__restrict__ apoly,
                  const fe *__restrict__ low_even, const fe *__restrict__ low_odd, const fe *__restrict__ hi, unsigned nhi,
                  const fe *__restrict__ binv, fe *__restrict__ out) {
    const unsigned kc = blockIdx.y, L = A->ncosets / 2;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + threadIdx.x;
    if (i >= n) return;
    const fe x = mul(A->shift[kc], W[i]);
    const fe *low = ((kc & 1) ? low_odd : low_even) + (unsigned long long)(kc >> 1) * n + i;
    acc192 s;
    for (unsigned g = 0; g < A->ngroups; g++)
        s.mac(mul(A->shift_adj[kc][g], W[(A->adj_mod[g] * i) & (n - 1)]), low[(unsigned long long)(1 + g) * L * n]);
    fe t = add(s.reduce(), low[0]);
    for (unsigned p = 0; p < nhi; p++) t = add(t, hi[((unsigned long long)p * A->ncosets + kc) * n + i]);
    fe res = mul(mul(t, sub(x, A->g_last)), A->zinv[kc]);
    const fe *cur = lde + A->lde_coset_stride[kc] + i;
    unsigned a = 0;
    for (unsigned g = 0; g < A->nbgroups; g++) {
        const fe xpb = mul(A->b_shift_adj[kc][g], W[(A->b_adj_mod[g] * i) & (n - 1)]);
        acc192 b;
        for (; a < A->nassertions && A->a_group[a] == g; a++) {
            fe v = A->a_value[a];
            if (A->a_poly_len[a] > 1) {
                const fe *poly = apoly + A->a_poly_off[a];
                const fe y = mul(x, A->a_xoff[a]);
                v = 0;
                for (unsigned m = A->a_poly_len[a]; m-- > 0;) v = add(mul(v, y), poly[m]);
            }
            b.mac(add(A->a_alpha[a], mul(A->a_beta[a], xpb)), sub(cur[(unsigned long long)A->a_col[a] * A->col_stride], v));
        }
        res = add(res, mul(b.reduce(), binv[((unsigned long long)g * A->ncosets + kc) * n + i]));
    }
    out[kc * n + i] = res;
}

// Low-degree split: the Rescue residuals and the linear rest have degree < (ce/2) * n for these AIRs, so their alpha and
// per-group beta parts are evaluated on the even ce cosets only (half the rows), interpolated there and evaluated on the
// odd cosets (NP * ce/2 size-n inverse transforms, an L x L mix, NP * ce/2 forward transforms); the curve items, whose
// degree needs the whole domain, run on every coset as before.
template <int AIR>
void launch_split(const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly, fe *part, fe *out, Stream &st,
                  cudaEvent_t *ev, const RootTable &rt, NttScratch &sc) {
    const unsigned long long n = 1ULL << h.logn;
    const unsigned gx = (unsigned)((n + CONS_THREADS - 1) / CONS_THREADS), ce = h.ncosets, L = ce / 2, NP = 1 + h.ngroups;
    constexpr int NR = airs::Items<AIR>::rescue, NE = airs::Items<AIR>::ecc;
    const size_t slab = (size_t)NP * L * n;
    fe *low_parts = part, *low_sum = low_parts + (size_t)(NR + 1) * slab, *low_coef = low_sum + slab, *low_mix = low_coef + slab,
       *low_odd = low_mix + slab, *hi = low_odd + slab, *binv = hi + (size_t)NE * ce * n;
    // interpolation on the even cosets, L x L mix of the per-coset coefficient sets, evaluation on the odd cosets, for `np`
    // polynomials laid out [p][L][n]
    const fe w2l = root_of_unity(ilog2(2 * L)), linv = inv(to_mont(L));
    std::vector<fe> mix(L * L);
    for (unsigned jo = 0; jo < L; jo++)
        for (unsigned je = 0; je < L; je++) {
            const unsigned e = (2 * (jo + L - je) + 1) % (2 * L);   // 2 (j' - j) + 1 mod 2L
            fe acc = 0, step = f63::pow(w2l, e), cur = ONE;
            for (unsigned t = 0; t < L; t++) { acc = add(acc, cur); cur = mul(cur, step); }
            mix[jo * L + je] = mul(acc, linv);
        }
    auto extend = [&](const fe *even, fe *coef, fe *mixed, fe *odd, unsigned np) {
        std::vector<fe> sinv(np * L), sodd(np * L);
        for (unsigned p = 0; p < np; p++)
            for (unsigned j = 0; j < L; j++) { sinv[p * L + j] = inv(h.shift[2 * j]); sodd[p * L + j] = h.shift[2 * j + 1]; }
        coset_intt_columns(rt, sc, even, n, coef, n, h.logn, sinv.data(), (size_t)np * L, st);
        coset_mix(coef, mixed, n, L, np, mix.data(), st);
        coset_ntt_entries(rt, sc, mixed, odd, (size_t)np * L, h.logn, sodd.data(), st);
        CSG_CUDA(cudaStreamSynchronize(st.s));   // the staging vectors are read by async copies
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
    const size_t eslab = (size_t)M.total * L * n;
    fe *eccl_even = binv + (size_t)CONS_MAX_BGROUPS * ce * n, *eccl_coef = eccl_even + eslab, *eccl_mix = eccl_coef + eslab, *eccl_odd = eccl_mix + eslab;
    auto mark = [&](int k) { if (ev) CSG_CUDA(cudaEventRecord(ev[k], st.s)); };
    mark(0);
    CSG_LAUNCH(st, (cons_low_kernel<AIR, 0>), dim3(gx, L, NR), CONS_THREADS, 0, args_dev, lde, W, ptab, low_parts, 0u);
    mark(1);
    if (ecc_split) {
        CSG_LAUNCH(st, cons_ecc_low_kernel<AIR>, dim3(gx, L, 4), CONS_THREADS, 0, args_dev, lde, ptab, M, eccl_even);
        extend(eccl_even, eccl_coef, eccl_mix, eccl_odd, M.total);
        CSG_LAUNCH(st, cons_ecc_merge_kernel<AIR>, dim3(gx, ce, 2), CONS_THREADS, 0, args_dev, lde, W, ptab, M, (const fe *)eccl_even, (const fe *)eccl_odd, hi);
    } else CSG_LAUNCH(st, (cons_item_kernel<AIR, 1>), dim3(gx, ce, 2), CONS_THREADS, 0, args_dev, lde, W, ptab, hi, 0u);
    mark(2);
    CSG_LAUNCH(st, (cons_item_kernel<AIR, 2>), dim3(gx, ce, 1), CONS_THREADS, 0, args_dev, lde, W, ptab, hi, 0u);
    mark(3);
    CSG_LAUNCH(st, (cons_low_kernel<AIR, 3>), dim3(gx, L, 1), CONS_THREADS, 0, args_dev, lde, W, ptab, low_parts, (unsigned)NR);
    sum_slices(low_parts, low_sum, slab, NR + 1, st);
    extend(low_sum, low_coef, low_mix, low_odd, NP);   // interpolate on the even cosets, evaluate on the odd ones
    if (h.nbgroups > 0)
        CSG_LAUNCH(st, boundary_inverse_kernel, dim3((unsigned)((n + INV_CHUNK * INV_THREADS - 1) / (INV_CHUNK * INV_THREADS)), ce, h.nbgroups),
                   INV_THREADS, 0, args_dev, W, binv);
    CSG_LAUNCH(st, cons_final_kernel, dim3(gx, ce), CONS_THREADS, 0, args_dev, lde, W, apoly, (const fe *)low_sum, (const fe *)low_odd, (const fe *)hi,
               (unsigned)NE, (const fe *)binv, out);
    mark(4);
}
}  // namespace

#if defined(CSG_REDC_CHECK)
unsigned long long redc_violations() {
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&v, f63::csg_redc_violations, sizeof v);
    return v;
}
#else
unsigned long long redc_violations() { return 0; }
#endif

size_t constraint_scratch_elements(int air_id, size_t n, size_t ncosets) {
    size_t items = 0;
    switch (air_id) {
    case airs::TRANSACTION: items = airs::Items<airs::TRANSACTION>::rescue + airs::Items<airs::TRANSACTION>::ecc; break;
    case airs::MERKLE_UPDATE: items = airs::Items<airs::MERKLE_UPDATE>::rescue; break;
    case airs::MERKLE_INIT: items = airs::Items<airs::MERKLE_INIT>::rescue; break;
    case airs::SCHNORR: items = airs::Items<airs::SCHNORR>::rescue + airs::Items<airs::SCHNORR>::ecc; break;
    case airs::RESCUE: items = airs::Items<airs::RESCUE>::rescue; break;
    default: break;
    }
    // split path: (rescue items + rest) x (1 + groups) x ce/2 partial slabs, 4 more slabs for the extension, curve items, divisors
    const size_t split = ((items + 1 + 4) * (1 + CONS_MAX_GROUPS) / 2 + 3 + CONS_MAX_BGROUPS + 4 * 16 / 2) * n * ncosets;   // + 4 slabs of <= 16 curve polynomials on half of the cosets
    const size_t plain = (items + CONS_MAX_BGROUPS) * n * ncosets;
    return split > plain ? split : plain;
}

void eval_constraints_ext2(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly,
                           fe *part, fe *out, Stream &st, cudaEvent_t *ev);   // constraints_ext.cu, compiled once per degree
void eval_constraints_ext3(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly,
                           fe *part, fe *out, Stream &st, cudaEvent_t *ev);
void eval_constraints_ext(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab,
                          const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev) {
    if (h.ext_degree == 2) eval_constraints_ext2(air_id, args_dev, h, lde, W, ptab, apoly, part, out, st, ev);
    else if (h.ext_degree == 3) eval_constraints_ext3(air_id, args_dev, h, lde, W, ptab, apoly, part, out, st, ev);
    else throw std::runtime_error("extension degree must be 2 or 3");
}

void eval_constraints(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab,
                      const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev, const RootTable *rt, NttScratch *sc) {
    const bool split = rt && sc && h.ncosets == 8 && h.ngroups <= (unsigned)airs::MAX_SPLIT_GROUPS;
    if (split && air_id == airs::TRANSACTION) { launch_split<airs::TRANSACTION>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev, *rt, *sc); return; }
    if (split && air_id == airs::SCHNORR) { launch_split<airs::SCHNORR>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev, *rt, *sc); return; }
    switch (air_id) {
    case airs::TRANSACTION: launch<airs::TRANSACTION>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::MERKLE_UPDATE: launch<airs::MERKLE_UPDATE>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::MERKLE_INIT: launch<airs::MERKLE_INIT>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::SCHNORR: launch<airs::SCHNORR>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::RANGE: launch<airs::RANGE>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::RESCUE: launch<airs::RESCUE>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    default: throw std::runtime_error("unknown AIR id");
    }
}

}  // namespace csg
