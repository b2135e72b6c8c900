// Constraint evaluation over the constraint-evaluation domain and merging into one composition column per coset
// (stage 3 of Prover::prove: winterfell ConstraintEvaluator::evaluate + the divisor step of
// ConstraintEvaluationTable::into_poly, driven from /root/reference/src/lib.rs:140 with the AIRs of src/air.rs & co).
//
// For every row x = s_k * w_n^i of every ce coset the kernel produces
//     C(x) = T(x) * (x - g^(n-1)) / (x^n - 1)  +  sum_groups  B_g(x) / (x^steps_g - offset_g)
// with T the merged transition constraints (airs.cuh) and B_g the merged boundary constraints of one divisor group.
// On a coset x^n is a constant, so the transition divisor costs one multiplication per row.
#pragma once
#include "dev.cuh"
#include "ntt.cuh"
#include "rescue_tables.h"

namespace csg {

constexpr int CONS_MAX_CONSTRAINTS = 128, CONS_MAX_GROUPS = 8, CONS_MAX_PERIODIC = 64, CONS_MAX_ASSERTIONS = 64,
              CONS_MAX_BGROUPS = 4, CONS_MAX_COSETS = 32;

// everything the kernel needs besides the trace; lives in device memory, read warp-uniformly
struct ConsArgs {
    // domain
    unsigned logn;                 // trace length n = 2^logn
    unsigned ncosets;              // ce cosets evaluated
    unsigned long long lde_coset_stride[CONS_MAX_COSETS];  // element offset of ce coset kc inside the LDE buffer
    unsigned long long col_stride;
    unsigned width;
    fe shift[CONS_MAX_COSETS];     // s_k
    fe zinv[CONS_MAX_COSETS];      // 1 / (s_k^n - 1)
    fe g_last;                     // g^(n-1)
    // transition combination
    unsigned nconstraints, ngroups;
    fe alpha[CONS_MAX_CONSTRAINTS], beta[CONS_MAX_CONSTRAINTS];
    unsigned char group[CONS_MAX_CONSTRAINTS];
    unsigned long long adj_mod[CONS_MAX_GROUPS];                 // adj_g mod n
    fe shift_adj[CONS_MAX_COSETS][CONS_MAX_GROUPS];              // s_k^adj_g
    // periodic columns: value of column c at row i of ce coset kc = ptab[kc * ptab_coset_stride + poff[c] + (i & pmask[c])]
    unsigned nperiodic;
    unsigned poff[CONS_MAX_PERIODIC], pmask[CONS_MAX_PERIODIC];
    unsigned long long ptab_coset_stride;
    // boundary constraints
    unsigned nbgroups, nassertions;
    unsigned long long b_adj_mod[CONS_MAX_BGROUPS], b_steps[CONS_MAX_BGROUPS];   // adj_g mod n, number of asserted steps
    fe b_offset[CONS_MAX_BGROUPS];                                               // g^(steps * first_step)
    fe b_shift_adj[CONS_MAX_COSETS][CONS_MAX_BGROUPS], b_shift_steps[CONS_MAX_COSETS][CONS_MAX_BGROUPS];  // s_k^adj, s_k^steps
    unsigned a_col[CONS_MAX_ASSERTIONS], a_group[CONS_MAX_ASSERTIONS];
    fe a_alpha[CONS_MAX_ASSERTIONS], a_beta[CONS_MAX_ASSERTIONS], a_value[CONS_MAX_ASSERTIONS];
    unsigned a_poly_len[CONS_MAX_ASSERTIONS];            // > 1: value polynomial of that many coefficients at a_poly_off
    unsigned long long a_poly_off[CONS_MAX_ASSERTIONS];
    fe a_xoff[CONS_MAX_ASSERTIONS];                      // the polynomial is evaluated at x * a_xoff
    // FieldExtension::Quadratic / Cubic: components 1 (and 2) of the E-valued coefficients; component 0 is alpha / beta above
    unsigned ext_degree;
    fe alpha_x[2][CONS_MAX_CONSTRAINTS], beta_x[2][CONS_MAX_CONSTRAINTS];
    fe a_alpha_x[2][CONS_MAX_ASSERTIONS], a_beta_x[2][CONS_MAX_ASSERTIONS];
    // base-field Rescue users: alpha / beta with the forward MDS product folded in (fill_rescue_tables)
    airs::RescueTables rt;
};
// after alpha, beta and group are set: the tables of every Rescue user of the AIR
void fill_rescue_tables(int air_id, ConsArgs &args);

// A sharded proof (comm.cuh) can still use the low-degree splits when every rank owns whole even/odd coset pairs: each rank
// interpolates on its own even cosets, the coefficient sets are all-gathered, and each rank mixes and evaluates its own odd cosets.
struct SplitExchange {
    unsigned world, rank;
    void (*gather)(void *self, void *buf, size_t bytes);   // in-place all-gather of equal slices on the proving stream
    void *self;
    fe *buf;                                              // world slices of the largest polynomial set
    size_t buf_elems;
};

// lde: coset-major extended trace; W: root table of size n; ptab / apoly: periodic tables and assertion value
// polynomials; part: scratch of constraint_scratch_elements() for the per-item partial sums; out[kc * n + i] receives C(x)
void eval_constraints(int air_id, const ConsArgs *args_dev, const ConsArgs &args_host, const fe *lde, const fe *W, const fe *ptab,
                      const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev = nullptr,   // ev: 6 events: [0..4] bracket the 4 phases, [5] follows the curve-formula kernel inside phase 2
                      const RootTable *rt = nullptr, NttScratch *sc = nullptr,   // given: low-degree constraints use half of the cosets (TX, Schnorr)
                      const SplitExchange *xch = nullptr);                      // given: args describe the cosets of one rank of a sharded proof
size_t constraint_scratch_elements(int air_id, size_t n, size_t ncosets);
// the same with E-valued coefficients (args.ext_degree = 2 or 3): one pass over the rows, component j of the merged column
// written to out[(j * ncosets + kc) * n + i]; part: ext_degree times the scratch of the base-field call
void eval_constraints_ext(int air_id, const ConsArgs *args_dev, const ConsArgs &args_host, const fe *lde, const fe *W, const fe *ptab,
                          const fe *apoly, fe *part, fe *out, Stream &st, cudaEvent_t *ev = nullptr,
                          const RootTable *rt = nullptr, NttScratch *sc = nullptr,   // given: the low-degree split, per component
                          const SplitExchange *xch = nullptr);

unsigned long long redc_violations();   // debug builds (-DCSG_REDC_CHECK): reductions entered with an out-of-range operand

}  // namespace csg
