// Coefficient tables of the Rescue-round constraints with the forward MDS product folded into the random linear combination.
//
// A Rescue residual (rescue::enforce_round, /root/reference/src/utils/rescue.rs:269-300) is
//     d_i = cube(sum_j INV_MDS[i][j] (next_j - ark_{14+j})) - ark_i - sum_j MDS[i][j] cube(cur_j),
// and a user of it (14 consecutive result slots s..s+13 under one flag) contributes sum_i c_{s+i} d_i to a part of T(x),
// c = alpha for the alpha part and beta (restricted to one degree group) for a beta part.  The last term is
//     sum_j (-sum_i c_{s+i} MDS[i][j]) cube(cur_j):
// 14 multiply-adds per part with a per-proof table instead of the 196 of the matrix product.  All arithmetic is exact
// modulo p, so the merged value is the one the reference computes.  Filled on the host once per proof
// (fill_rescue_tables, constraints.cu); `use` = 2 * (Rescue item of the AIR) + (0: first user, 1: second user).
#pragma once
#include <stdint.h>

namespace airs {

constexpr int RT_MAX_USES = 10, RT_MAX_GROUPS = 3;

struct RescueTables {
    unsigned char ng[RT_MAX_USES];                    // distinct degree groups among the 14 slots of the use
    unsigned char grp[RT_MAX_USES][RT_MAX_GROUPS];    // those groups
    uint64_t a_bwd[RT_MAX_USES][14];                  // alpha_{s+i}
    uint64_t b_bwd[RT_MAX_USES][RT_MAX_GROUPS][14];   // beta_{s+i} where slot s+i is in group grp[q], else 0
    uint64_t a_fwd[RT_MAX_USES][14];                  // -sum_i alpha_{s+i} MDS[i][j]
    uint64_t b_fwd[RT_MAX_USES][RT_MAX_GROUPS][14];   // -sum_{i in group grp[q]} beta_{s+i} MDS[i][j]
};

}  // namespace airs
