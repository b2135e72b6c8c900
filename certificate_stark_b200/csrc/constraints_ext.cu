// Constraint evaluation with E-valued random coefficients (FieldExtension::Quadratic / Cubic): the combined-mode kernels of
// constraints_kernels.cuh instantiated with CSG_EXT_DEG accumulators per thread.  Compiled once per degree (Makefile) so the
// two sets of instantiations build in parallel.
#include <stdexcept>

#include "constraints_kernels.cuh"

#ifndef CSG_EXT_DEG
#error "compile with -DCSG_EXT_DEG=2 or 3"
#endif
#define CSG_CAT2(a, b) a##b
#define CSG_CAT(a, b) CSG_CAT2(a, b)

namespace csg {

void CSG_CAT(eval_constraints_ext, CSG_EXT_DEG)(int air_id, const ConsArgs *args_dev, const ConsArgs &h, const fe *lde, const fe *W, const fe *ptab,
                                                const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev, const RootTable *rt, NttScratch *sc, const SplitExchange *xch) {
    constexpr int D = CSG_EXT_DEG;
    // the low-degree split (rescue residuals, linear rest, curve formulas on the even cosets only), per component
    const bool split = rt && sc && split_applies(h, xch);
    if (split && air_id == airs::TRANSACTION) { launch_split<airs::TRANSACTION, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev, *rt, *sc, xch); return; }
    if (split && air_id == airs::SCHNORR) { launch_split<airs::SCHNORR, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev, *rt, *sc, xch); return; }
    switch (air_id) {
    case airs::TRANSACTION: launch<airs::TRANSACTION, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::MERKLE_UPDATE: launch<airs::MERKLE_UPDATE, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::MERKLE_INIT: launch<airs::MERKLE_INIT, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::SCHNORR: launch<airs::SCHNORR, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::RANGE: launch<airs::RANGE, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    case airs::RESCUE: launch<airs::RESCUE, D>(args_dev, h, lde, W, ptab, apoly, part, out, st, ev); break;
    default: throw std::runtime_error("unknown AIR id");
    }
}

}  // namespace csg
