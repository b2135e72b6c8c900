// See constraints.cuh / airs.cuh.  Columns are read coalesced along the row index.
//
// Launch structure per proof: one `cons_item_kernel` over (rows, ce cosets, heavy items) -- a thread evaluates ONE Rescue
// residual or ONE curve item of one row and writes its partial sum of T(x) -- then `cons_rest_kernel` over (rows, cosets)
// adds the cheap linear constraints, sums the partials, applies the divisors and the boundary constraints and writes the
// merged column.  Splitting the row's work this way keeps each kernel's code and register footprint small (the
// monolithic version ran at 8 warps/SM with 20 % of its stalls on instruction fetch) and multiplies the parallelism.
#include <cstring>
#include <stdexcept>
#include <vector>

#include "airs.cuh"
#include "constraints.cuh"
#include "constraints_kernels.cuh"
#include "stages.cuh"

namespace csg {
using namespace f63;

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

void fill_rescue_tables(int air_id, ConsArgs &h) { airs::fill_rescue_tables(air_id, h.alpha, h.beta, h.group, h.nconstraints, h.rt); }

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
                           fe *part, fe *out, Stream &st, cudaEvent_t *ev, const RootTable *rt, NttScratch *sc, const SplitExchange *xch);   // constraints_ext.cu, compiled once per degree
void eval_constraints_ext3(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab, const fe *apoly,
                           fe *part, fe *out, Stream &st, cudaEvent_t *ev, const RootTable *rt, NttScratch *sc, const SplitExchange *xch);
void eval_constraints_ext(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab,
                          const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev, const RootTable *rt, NttScratch *sc, const SplitExchange *xch) {
    if (h.ext_degree == 2) eval_constraints_ext2(air_id, args_dev, h, lde, W, ptab, apoly, part, out, st, ev, rt, sc, xch);
    else if (h.ext_degree == 3) eval_constraints_ext3(air_id, args_dev, h, lde, W, ptab, apoly, part, out, st, ev, rt, sc, xch);
    else throw std::runtime_error("extension degree must be 2 or 3");
}

void eval_constraints(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab,
                      const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev, const RootTable *rt, NttScratch *sc, const SplitExchange *xch) {
    const bool split = rt && sc && split_applies(h, xch);
    if (split && air_id == airs::TRANSACTION) { launch_split<airs::TRANSACTION>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev, *rt, *sc, xch); return; }
    if (split && air_id == airs::SCHNORR) { launch_split<airs::SCHNORR>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev, *rt, *sc, xch); return; }
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
