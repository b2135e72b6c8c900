// See constraints.cuh / airs.cuh.  One thread = one row of one ce coset; columns are read coalesced along the row index.
#include "airs.cuh"
#include "constraints.cuh"

namespace csg {
using namespace f63;

namespace {
constexpr int CONS_THREADS = 128;

template <int AIR>
__global__ void __launch_bounds__(CONS_THREADS) cons_kernel(const ConsArgs *__restrict__ A, const fe *__restrict__ lde, const fe *__restrict__ W,
                                                            const fe *__restrict__ ptab, const fe *__restrict__ apoly, fe *__restrict__ out) {
    __shared__ fe xp_s[airs::MAX_GROUPS][CONS_THREADS];   // x^adj of each degree group, one column per thread
    const unsigned tid = threadIdx.x, kc = blockIdx.y;
    const unsigned long long n = 1ULL << A->logn, i = blockIdx.x * (unsigned long long)CONS_THREADS + tid;
    if (i >= n) return;
    const unsigned long long inext = (i + 1) & (n - 1);
    const fe *base = lde + A->lde_coset_stride[kc];
    airs::Frame f{base + i, base + inext, (size_t)A->col_stride};
    airs::Periodic pv{ptab + kc * A->ptab_coset_stride, A->poff, A->pmask, (uint32_t)i};

    const fe x = mul(A->shift[kc], W[i]);
    for (unsigned g = 0; g < A->ngroups; g++) xp_s[g][tid] = mul(A->shift_adj[kc][g], W[(A->adj_mod[g] * i) & (n - 1)]);
    airs::Comb C{A->alpha, A->beta, A->group, &xp_s[0][tid], (size_t)CONS_THREADS, acc192()};
    airs::eval_transition<AIR>(f, pv, C);
    fe res = mul(mul(C.sum.reduce(), sub(x, A->g_last)), A->zinv[kc]);

    unsigned a = 0;
    for (unsigned g = 0; g < A->nbgroups; g++) {
        const fe xpb = mul(A->b_shift_adj[kc][g], W[(A->b_adj_mod[g] * i) & (n - 1)]);
        acc192 s;
        for (; a < A->nassertions && A->a_group[a] == g; a++) {
            fe v = A->a_value[a];
            if (A->a_poly_len[a] > 1) {   // Assertion::sequence: value polynomial evaluated at x * g^-first_step
                const fe *poly = apoly + A->a_poly_off[a];
                const fe y = mul(x, A->a_xoff[a]);
                v = 0;
                for (unsigned m = A->a_poly_len[a]; m-- > 0;) v = add(mul(v, y), poly[m]);
            }
            s.mac(add(A->a_alpha[a], mul(A->a_beta[a], xpb)), sub(f.cur(A->a_col[a]), v));
        }
        const fe xs = A->b_steps[g] == 1 ? x : mul(A->b_shift_steps[kc][g], W[(A->b_steps[g] * i) & (n - 1)]);
        res = add(res, mul(s.reduce(), inv(sub(xs, A->b_offset[g]))));
    }
    out[kc * n + i] = res;
}

template <int AIR>
void launch(const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly, fe *out, Stream &st) {
    const unsigned long long n = 1ULL << h.logn;
    dim3 grid((unsigned)((n + CONS_THREADS - 1) / CONS_THREADS), h.ncosets);
    CSG_LAUNCH(st, cons_kernel<AIR>, grid, CONS_THREADS, 0, args_dev, lde, W, ptab, apoly, out);
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

void eval_constraints(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab,
                      const fe *apoly, fe *out, Stream &st) {
    switch (air_id) {
    case airs::TRANSACTION: launch<airs::TRANSACTION>(args_dev, h, lde, W, ptab, apoly, out, st); break;
    case airs::MERKLE_UPDATE: launch<airs::MERKLE_UPDATE>(args_dev, h, lde, W, ptab, apoly, out, st); break;
    case airs::MERKLE_INIT: launch<airs::MERKLE_INIT>(args_dev, h, lde, W, ptab, apoly, out, st); break;
    case airs::SCHNORR: launch<airs::SCHNORR>(args_dev, h, lde, W, ptab, apoly, out, st); break;
    case airs::RANGE: launch<airs::RANGE>(args_dev, h, lde, W, ptab, apoly, out, st); break;
    case airs::RESCUE: launch<airs::RESCUE>(args_dev, h, lde, W, ptab, apoly, out, st); break;
    default: throw std::runtime_error("unknown AIR id");
    }
}

}  // namespace csg
