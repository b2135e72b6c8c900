// Per-row transition-constraint evaluation of the six AIRs of the reference, fused with winterfell's random linear
// combination (K5 of SURVEY.md section 2.2).  What the reference computes per constraint-evaluation-domain row is
//     result[] = 0;  Air::evaluate_transition(frame, periodic_values, result);      (src/air.rs:114-173 -> :383-610)
//     T(x) = sum_i result[i] * (alpha_i + beta_i * x^adj(group(i)))                    (winterfell ConstraintEvaluator)
// Here result[] is never materialised: every contribution `result[slot] += flag * value` (agg_constraint,
// src/utils/mod.rs:58-62) is folded straight into the 192-bit lazily reduced accumulator of T, grouped by flag so
// that a flag is multiplied in once per group.  All arithmetic is exact modulo p, so the value of T is identical to
// the reference's whatever the order of evaluation.
//
// Reference functions restated (polynomial identities only, no code shared):
//   TransactionAir::evaluate_transition / evaluate_constraints      src/air.rs:114-173, 383-610
//   merkle::init::evaluate_constraints                              src/merkle/init/air.rs:159-202
//   merkle::update::evaluate_constraints / evaluate_merkle_update_auth   src/merkle/update/air.rs:215-369
//   schnorr::evaluate_constraints / enforce_hash_copy               src/schnorr/air.rs:394-531, 309-330
//   ecc::enforce_point_doubling / _addition_mixed / _addition_reduce_x   src/utils/ecc.rs:73-172
//   rescue::enforce_round                                           src/utils/rescue.rs:269-300
//   field::enforce_double_and_add_step(_constrained)                src/utils/field.rs:31-70
//   RangeProofAir / RescueAir evaluate_transition                   src/range/air.rs:69-105, benches/rescue.rs:205-222
#pragma once
#include "ecc.cuh"
#include "rescue.cuh"
#include "rescue_tables.h"
#include <stdexcept>

// Unroll factor of the short per-element loops of the linear constraints (device).  Fully unrolled, the linear-rest kernel is
// 150 KB of straight-line code that every warp streams through once, and ncu attributes 30 % of its stall cycles to
// instruction fetch (no_instruction); rolled loops measured 0.2 ms faster.  (Making the accumulator reduction a real call
// halves the code again but measured 0.9 ms slower.)
#if defined(__CUDA_ARCH__)
#define CSG_PRAGMA_(x) _Pragma(#x)
#define CSG_PRAGMA(x) CSG_PRAGMA_(x)
#else
#define CSG_PRAGMA(x)
#endif
#ifndef CSG_REST_UNROLL
#define CSG_REST_UNROLL 1
#endif
#if defined(__CUDA_ARCH__)
#define CSG_REST_LOOP CSG_PRAGMA(unroll CSG_REST_UNROLL)
#else
#define CSG_REST_LOOP
#endif
#ifndef CSG_RESCUE_FWD_UNROLL
#define CSG_RESCUE_FWD_UNROLL 2
#endif

namespace airs {
using f63::fe;

enum : int { TRANSACTION = 0, MERKLE_UPDATE = 1, MERKLE_INIT = 2, SCHNORR = 3, RANGE = 4, RESCUE = 5 };

// ---- layout (src/merkle/constants.rs:27-56, src/constants.rs:35-116, src/schnorr/constants.rs) ----
enum : int {
    HSW = 14, HRW = 7, APW = 12, PPW = 18, PCW = 6,
    SENDER_INITIAL = 0, SENDER_BIT = 14, SENDER_UPDATED = 15, RECEIVER_INITIAL = 29, RECEIVER_BIT = 43, RECEIVER_UPDATED = 44,
    PREV_ROOT = 58, VALUE_RES = 65, BALANCE_RES = 90, NONCE_RES = 91, INT_ROOT_RES = 92, ROOT_MATCH_RES = 99,
    SENDER_KEY = 65, RECEIVER_KEY = 77, DELTA_COPY = 89, SIGMA_COPY = 90, NONCE_COPY = 91,
    SENDER_KEY_RES = 101, RECEIVER_KEY_RES = 103, DELTA_COPY_RES = 105, SIGMA_COPY_RES = 106, NONCE_COPY_RES = 107,
    DELTA_RANGE_RES = 108, SIGMA_RANGE_RES = 109,
    DELTA_BIT = 56, DELTA_ACC = 57, SIGMA_BIT = 92, SIGMA_ACC = 93,
    LIMBS = 2 * PPW + 1,   // 37: h bit, then the four limb accumulators
    SIG_HASH = 2 * PPW + 6  // 42: Rescue state of the message hash
};
// periodic-column indices of the transaction AIR (src/constants.rs:85-116)
enum : int {
    TX_SETUP = 0, TX_MERKLE = 1, TX_HASH_INPUT = 2, TX_FINISH = 3, TX_HASH = 4, TX_SCHNORR = 5, TX_SCALAR_MULT = 6, TX_DOUBLING = 7,
    TX_DIGEST = 8, TX_SCHNORR_HASH = 12, TX_INTERNAL = 13, TX_RANGE_STEP = 17, TX_RANGE_FINISH = 18, TX_VALUE_COPY = 19, TX_ARK = 20
};

constexpr int MAX_GROUPS = 8;

// two consecutive rows of the (extended) trace; column c of the current row is cur_p[c * stride]
struct Frame {
    const fe *cur_p, *next_p;
    size_t stride;
    CSG_HD fe cur(int c) const { return cur_p[(size_t)c * stride]; }
    CSG_HD fe next(int c) const { return next_p[(size_t)c * stride]; }
};
// periodic column c at this row: tab[off[c] + (i & mask[c])]
struct Periodic {
    const fe *tab;
    const uint32_t *off, *mask;
    uint32_t i;
    CSG_HD fe operator()(int c) const { return tab[off[c] + (i & mask[c])]; }
};

// alpha[s] + beta[s] * xp[group[s]].  A real call on the device: it is used at several hundred unrolled sites and inlining
// it there made the instruction footprint of the kernels exceed the instruction cache.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
inline
#endif
fe slot_coefficient(const fe *alpha, const fe *beta, const uint8_t *group, const fe *xp, size_t xp_stride, int slot) {
    return f63::add(alpha[slot], f63::mul(beta[slot], xp[group[slot] * xp_stride]));
}

// The random linear combination T(x) = sum_s result_s * (alpha_s + beta_s * x^adj(group(s))).
//   combined mode (SPLIT = false): `sum` accumulates T(x) itself, the coefficient of a slot is alpha_s + beta_s * xp[g].
//   split mode    (SPLIT = true):  `sum` accumulates A = sum_s alpha_s * result_s and part[g] accumulates
//                                  B_g = sum_{s in g} beta_s * result_s, so that T = A + sum_g x^adj_g * B_g can be formed later.
// Split mode exists for the constraints whose degree stays below half the composition degree (Rescue rounds, Merkle and
// copy logic: < 4n for the transaction and Schnorr AIRs): A and the B_g are then polynomials of degree < 4n, evaluated on
// half of the constraint-evaluation cosets only and extended to the other half with two small transforms.
// Extension fields (DEG = 2, 3): the coefficients alpha_s, beta_s are elements of E while the constraint values stay in the
// base field, so component j of T(x) is the same combination with component j of every coefficient.  `sum` accumulates
// component 0 and sum_x[j-1] component j; the extra components live at alpha_x[(j-1) * coef_stride + slot].  One evaluation
// of the constraints feeds all DEG accumulators (combined mode only).
constexpr int MAX_SPLIT_GROUPS = 6;
template <bool SPLIT, int DEG = 1>
struct CombT {
    static constexpr bool split = SPLIT;
    static constexpr int D = DEG;

    const fe *alpha, *beta;
    const uint8_t *group;
    const fe *xp;        // x^adj per degree group, element g at xp[g * xp_stride]   (combined mode only)
    size_t xp_stride;
    f63::acc192 sum;
    // split mode: the 192-bit accumulator of B_g lives in memory (shared memory on the device), word k at
    // part[(3 * g + k) * part_stride] -- the group of a slot is only known at run time, and a register file cannot be indexed
    uint64_t *part;
    size_t part_stride;
    const fe *alpha_x = nullptr, *beta_x = nullptr;
    size_t coef_stride = 0;
    f63::acc192 sum_x[DEG > 1 ? DEG - 1 : 1];
    const RescueTables *rt = nullptr;   // device, base field: the Rescue users' coefficient tables (rescue_tables.h)
    CSG_HD fe coef(int slot) const { return slot_coefficient(alpha, beta, group, xp, xp_stride, slot); }
    CSG_HD fe coef_x(int j, int slot) const { return slot_coefficient(alpha_x + (j - 1) * coef_stride, beta_x + (j - 1) * coef_stride, group, xp, xp_stride, slot); }
    CSG_HD void add(int slot, fe v) {
        if (!SPLIT) {
            sum.mac(coef(slot), v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 1; j < DEG; j++) sum_x[j - 1].mac(coef_x(j, slot), v);
            return;
        }
        const int g = group[slot];
        sum.mac(alpha[slot], v);
        part_mac(0, g, beta[slot], v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 1; j < DEG; j++) {
            sum_x[j - 1].mac(alpha_x[(j - 1) * coef_stride + slot], v);
            part_mac(j, g, beta_x[(j - 1) * coef_stride + slot], v);
        }
    }
    // the accumulator of B_g of component j: words at part[((j * MAX_SPLIT_GROUPS + g) * 3 + k) * part_stride]
    CSG_HD void part_mac(int j, int g, fe b, fe v) {
        uint64_t *q = part + (size_t)(j * MAX_SPLIT_GROUPS + g) * 3 * part_stride;
        f63::acc192w t;
        t.lo = q[0]; t.mid = q[part_stride]; t.hi = q[2 * part_stride];
        t.mac(b, v);
        q[0] = t.lo; q[part_stride] = t.mid; q[2 * part_stride] = t.hi;
    }
    // Split mode on the device: a CTA barrier.  The linear-rest code is 130 KB of straight-line instructions that every warp
    // walks once per row tile, and instruction fetch was its top stall (ncu: no_instruction); a barrier every few hundred
    // instructions keeps the four warps of a CTA inside the same stretch, so one instruction-cache fill serves all of them:
    // cons_rest 6.97 -> 6.53 ms at 2^20 rows (gpurun_out/ab_rsync*.txt; one more barrier per loop iteration measured slower).
    // Every thread of the CTA reaches it: the split kernels have no divergent exit after their bounds check.
    CSG_HD void sync() const {
#if defined(__CUDA_ARCH__) && !defined(CSG_NO_REST_SYNC)
        if (SPLIT) __syncthreads();
#endif
    }
    CSG_HD fe part_value(int g, int j = 0) const {   // split mode: B_g of component j, reduced
        const uint64_t *q = part + (size_t)(j * MAX_SPLIT_GROUPS + g) * 3 * part_stride;
        f63::acc192w t;
        t.lo = q[0]; t.mid = q[part_stride]; t.hi = q[2 * part_stride];
        return t.reduce();
    }
};
using Comb = CombT<false>;
using SplitComb = CombT<true>;
// contributions that share one flag.  Combined mode: sum_k coef(slot_k) * v_k, multiplied by the flag once at flush time.
// Split mode: the alpha part likewise; the beta part of the current degree group is accumulated in registers and handed
// to B_g (times the flag) whenever the group of the next slot differs -- slots come in runs of equal declared degree, so
// that is rare, and no value is multiplied by the flag on its own.
struct FlagAcc {
    f63::acc192 s, sx[2];   // sx: components 1, 2 of E-valued coefficients (untouched, hence free, in the base field)
    f63::acc192 sb, sbx[2]; // split mode: beta part of group `gcur`
    int gcur = -1;
    fe flag;
    CSG_HD explicit FlagAcc(fe f) : flag(f) {}
    template <class CB> CSG_HD void add(CB &C, int slot, fe v) {
        if (!CB::split) {
            s.mac(C.coef(slot), v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 1; j < CB::D; j++) sx[j - 1].mac(C.coef_x(j, slot), v);
        } else {
            const int g = C.group[slot];
            if (g != gcur) { flush_beta(C); gcur = g; }
            s.mac(C.alpha[slot], v);
            sb.mac(C.beta[slot], v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 1; j < CB::D; j++) {
                sx[j - 1].mac(C.alpha_x[(j - 1) * C.coef_stride + slot], v);
                sbx[j - 1].mac(C.beta_x[(j - 1) * C.coef_stride + slot], v);
            }
        }
    }
    template <class CB> CSG_HD void flush_beta(CB &C) {
        if (gcur < 0) return;
        C.part_mac(0, gcur, flag, sb.reduce());
        sb = f63::acc192();
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 1; j < CB::D; j++) { C.part_mac(j, gcur, flag, sbx[j - 1].reduce()); sbx[j - 1] = f63::acc192(); }
    }
    template <class CB> CSG_HD void flush(CB &C) {
        if (CB::split) flush_beta(C);
        C.sum.mac(flag, s.reduce());
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 1; j < CB::D; j++) C.sum_x[j - 1].mac(flag, sx[j - 1].reduce());
    }
};

CSG_HD fe f_not(fe a) { return f63::sub(f63::ONE, a); }
CSG_HD fe f_bin(fe a) { return f63::sub(f63::sqr(a), a); }

// ---- Rescue round: backward_half(next) - forward_half(cur) of the 14-wide state at column col0, round constants from
// periodic columns ark0..ark0+27; up to two (flag, first slot) users of the same residual.
//   forward_half(cur)[i]  = sum_j MDS[i][j] * cur[j]^3 + ark[i]                      (src/utils/rescue.rs:274-279)
//   backward_half(next)[i] = (sum_j INV_MDS[i][j] * (next[j] - ark[14+j]))^3          (src/utils/rescue.rs:281-287)
// Row i of both products is formed, cubed/offset and consumed inside one rolled loop, so no 14-element intermediate is
// ever indexed dynamically (on the GPU that would put it in local memory: the first version of this kernel wrote 4x its
// algorithmic bytes to DRAM that way).
template <class PV, class CB>
CSG_HD void rescue_state(const Frame &f, const PV &pv, CB &C, int col0, int ark0, fe flag_a, int slot_a, bool second, fe flag_b, int slot_b) {
    fe tc[14], tn[14];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 14; j++) {
        tc[j] = rescue::cube(f.cur(col0 + j));
        tn[j] = f63::sub(f.next(col0 + j), pv(ark0 + 14 + j));
    }
    FlagAcc a(flag_a), b(flag_b);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 14; i++) {
        const uint64_t *mds = CSG_TABLE(CSG_MDS) + i * 14, *inv_mds = CSG_TABLE(CSG_INV_MDS) + i * 14;
        f63::acc128 fwd, bwd;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 14; j++) { fwd.mac(mds[j], tc[j]); bwd.mac(inv_mds[j], tn[j]); }
        const fe d = f63::sub(rescue::cube(bwd.reduce()), f63::add(fwd.reduce(), pv(ark0 + i)));
        a.add(C, slot_a + i, d);
        if (second) b.add(C, slot_b + i, d);
    }
    a.flush(C);
    if (second) b.flush(C);
}

// ---- merkle::update::evaluate_merkle_update_auth without its two Rescue rounds (src/merkle/update/air.rs:291-369)
template <class CB>
CSG_HD void merkle_auth_path(const Frame &f, CB &C, int base, fe tx_hash, fe hash_input, fe hashf) {
    const fe copy_flag = f63::mul(tx_hash, f_not(f63::add(hashf, hash_input)));
    const fe init_flag = f63::mul(tx_hash, hash_input);
    const fe bit = f.next(base + HSW), nbit = f_not(bit);
    C.add(base + HSW, f63::mul(tx_hash, f_bin(bit)));
    const fe init_bit = f63::mul(init_flag, bit), init_nbit = f63::mul(init_flag, nbit);
    FlagAcc keep(f63::add(copy_flag, init_nbit)), to_rate(init_bit), place_bit(init_bit), place_nbit(init_nbit);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < 2; k++) {
        const int o = base + k * (HSW + 1);
        CSG_REST_LOOP
        for (int i = 0; i < HRW; i++) {
            fe c = f.cur(o + i);
            keep.add(C, o + i, f63::sub(c, f.next(o + i)));                 // copy_flag and init_flag*(1-bit) share this difference
            to_rate.add(C, o + HRW + i, f63::sub(c, f.next(o + HRW + i)));  // init_flag*bit: the hash moves to the rate half
        }
        C.sync();
    }
    CSG_REST_LOOP
    for (int i = 0; i < HRW; i++) place_bit.add(C, base + i, f63::sub(f.next(base + HSW + 1 + i), f.next(base + i)));
    CSG_REST_LOOP
    for (int i = HRW; i < HSW; i++) place_nbit.add(C, base + i, f63::sub(f.next(base + HSW + 1 + i), f.next(base + i)));
    C.sync();
    keep.flush(C);
    to_rate.flush(C);
    place_bit.flush(C);
    place_nbit.flush(C);
}
// root bookkeeping of merkle::update::evaluate_constraints (src/merkle/update/air.rs:250-288)
template <class CB>
CSG_HD void merkle_roots(const Frame &f, CB &C, fe finish) {
    FlagAcc carry(f_not(finish)), fin(finish);
    CSG_REST_LOOP
    for (int i = 0; i < HRW; i++) {
        fe nr = f.next(PREV_ROOT + i), cr = f.cur(PREV_ROOT + i);
        carry.add(C, PREV_ROOT + i, f63::sub(nr, cr));
        fin.add(C, PREV_ROOT + i, f63::sub(nr, f.next(RECEIVER_UPDATED + i)));
        fin.add(C, INT_ROOT_RES + i, f63::sub(f.cur(SENDER_UPDATED + i), f.cur(RECEIVER_INITIAL + i)));
        fin.add(C, ROOT_MATCH_RES + i, f63::sub(f.next(SENDER_INITIAL + i), cr));
    }
    carry.flush(C);
    fin.flush(C);
}
// value / balance / nonce block (src/merkle/update/air.rs:96-144 and src/air.rs:405-453), added to `a`
template <class CB>
CSG_HD void value_block(const Frame &f, CB &C, FlagAcc &a) {
    CSG_REST_LOOP
    for (int i = 0; i < APW; i++) {
        a.add(C, VALUE_RES + i, f63::sub(f.cur(SENDER_INITIAL + i), f.cur(SENDER_UPDATED + i)));
        a.add(C, VALUE_RES + APW + i, f63::sub(f.cur(RECEIVER_INITIAL + i), f.cur(RECEIVER_UPDATED + i)));
    }
    a.add(C, VALUE_RES + 2 * APW, f63::sub(f.cur(RECEIVER_INITIAL + APW + 1), f.cur(RECEIVER_UPDATED + APW + 1)));
    a.add(C, BALANCE_RES, f63::sub(f63::sub(f.cur(SENDER_INITIAL + APW), f.cur(SENDER_UPDATED + APW)),
                                   f63::sub(f.cur(RECEIVER_UPDATED + APW), f.cur(RECEIVER_INITIAL + APW))));
    a.add(C, NONCE_RES, f63::sub(f.cur(SENDER_UPDATED + APW + 1), f63::add(f.cur(SENDER_INITIAL + APW + 1), f63::ONE)));
}

// ---- curve part of schnorr::evaluate_constraints: one scalar multiplication register bank (point at column o, its
// bit at o+18) against the affine point q (src/schnorr/air.rs:415-452, src/utils/ecc.rs:73-144)
template <class CB>
CSG_HD void scalar_mult_bank(const Frame &f, CB &C, int o, const fe (&q)[12], fe doubling, fe addition) {
    // operands are (re)loaded where they are used instead of being held across the two formulas: the curve arithmetic
    // alone needs ~100 64-bit temporaries, and the loads hit L1
    auto load_point = [&]() {
        ecc::point p;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 6; i++) { p.x.c[i] = f.cur(o + i); p.y.c[i] = f.cur(o + 6 + i); p.z.c[i] = f.cur(o + 12 + i); }
        return p;
    };
    const fe bit = f.cur(o + PPW);
    {
        ecc::point d = ecc::double_point(load_point());
        FlagAcc a(doubling);
        for (int i = 0; i < 6; i++) {
            a.add(C, o + i, f63::sub(f.next(o + i), d.x.c[i]));
            a.add(C, o + 6 + i, f63::sub(f.next(o + 6 + i), d.y.c[i]));
            a.add(C, o + 12 + i, f63::sub(f.next(o + 12 + i), d.z.c[i]));
        }
        a.add(C, o + PPW, f_bin(bit));
        a.flush(C);
    }
    {
        ecc::point m = ecc::add_mixed(load_point(), ecc::load6(q), ecc::load6(q + 6));
        const fe nbit = f_not(bit);
        FlagAcc a(addition);
        for (int i = 0; i < 6; i++) {
            a.add(C, o + i, f63::sub(f.next(o + i), f63::add(f63::mul(bit, m.x.c[i]), f63::mul(nbit, f.cur(o + i)))));
            a.add(C, o + 6 + i, f63::sub(f.next(o + 6 + i), f63::add(f63::mul(bit, m.y.c[i]), f63::mul(nbit, f.cur(o + 6 + i)))));
            a.add(C, o + 12 + i, f63::sub(f.next(o + 12 + i), f63::add(f63::mul(bit, m.z.c[i]), f63::mul(nbit, f.cur(o + 12 + i)))));
        }
        a.add(C, o + PPW, f63::sub(bit, f.next(o + PPW)));
        a.flush(C);
    }
}
// ---- the same bank for the low-degree split of the curve constraints.  The outputs of either formula are polynomials of
// degree 4(n-1) in the trace polynomials; only their products with the row's bit and the periodic flags exceed 4n.  So
//     T_bank = doubling * (sum_i c_i next_i + c_b bit(bit-1) - Cd)  +  addition * (sum_i c_i (next_i - (1-bit) cur_i) + c_b (bit - bit') - bit * Cm)
// with Cd = sum_i c_i d_i, Cm = sum_i c_i m_i the merged outputs of the doubling / mixed-addition formula: Cd and Cm (alpha
// part and per-group beta parts, split mode) are evaluated on the even cosets only and extended by NTT, the rest is cheap.
template <class CB>
CSG_HD void scalar_mult_bank_outputs(const Frame &f, CB &C, int o, const fe (&q)[12], int formula) {
    ecc::point p;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; i++) { p.x.c[i] = f.cur(o + i); p.y.c[i] = f.cur(o + 6 + i); p.z.c[i] = f.cur(o + 12 + i); }
    const ecc::point r = formula == 0 ? ecc::double_point(p) : ecc::add_mixed(p, ecc::load6(q), ecc::load6(q + 6));
    // slot order = column order: the runs of equal degree group stay together (x, then y and z)
    FlagAcc a(f63::ONE);
    for (int i = 0; i < 6; i++) a.add(C, o + i, r.x.c[i]);
    for (int i = 0; i < 6; i++) a.add(C, o + 6 + i, r.y.c[i]);
    for (int i = 0; i < 6; i++) a.add(C, o + 12 + i, r.z.c[i]);
    a.flush(C);
}
template <class CB>
CSG_HD fe scalar_mult_bank_merge(const Frame &f, const CB &C, int comp, int o, fe doubling, fe addition, fe Cd, fe Cm) {   // comp: component of E-valued coefficients
    const fe bit = f.cur(o + PPW), nbit = f_not(bit);
    f63::acc192 sd, sa;
    for (int i = 0; i < PPW; i++) {
        const fe c = comp ? C.coef_x(comp, o + i) : C.coef(o + i), nx = f.next(o + i);
        sd.mac(c, nx);
        sa.mac(c, f63::sub(nx, f63::mul(nbit, f.cur(o + i))));
    }
    const fe cb = comp ? C.coef_x(comp, o + PPW) : C.coef(o + PPW);
    const fe td = f63::sub(f63::add(sd.reduce(), f63::mul(cb, f_bin(bit))), Cd);
    const fe ta = f63::sub(f63::add(sa.reduce(), f63::mul(cb, f63::sub(bit, f.next(o + PPW)))), f63::mul(bit, Cm));
    return f63::add(f63::mul(doubling, td), f63::mul(addition, ta));
}
// last step of a signature: S + h.P, x reduced to affine, and h must equal the hash output
// (src/schnorr/air.rs:506-530, src/utils/ecc.rs:146-172)
template <class CB>
CSG_HD void schnorr_final_addition(const Frame &f, CB &C, fe final_add) {
    ecc::point s, hp;
    for (int i = 0; i < 6; i++) {
        s.x.c[i] = f.cur(i); s.y.c[i] = f.cur(6 + i); s.z.c[i] = f.cur(12 + i);
        hp.x.c[i] = f.cur(PPW + 1 + i); hp.y.c[i] = f.cur(PPW + 7 + i); hp.z.c[i] = f.cur(PPW + 13 + i);
    }
    ecc::point r = ecc::add_full(s, hp);
    ecc::fp6 nx;
    for (int i = 0; i < 6; i++) nx.c[i] = f.next(i);
    ecc::fp6 xz = ecc::mul(nx, r.z);
    FlagAcc a(final_add);
    for (int i = 0; i < 6; i++) {
        a.add(C, i, f63::sub(xz.c[i], r.x.c[i]));
        a.add(C, 6 + i, f63::sub(f.next(6 + i), r.y.c[i]));
        a.add(C, 12 + i, f63::sub(f.next(12 + i), r.z.c[i]));
    }
    for (int i = 0; i < 4; i++) a.add(C, LIMBS + 1 + i, f63::sub(f.cur(LIMBS + 1 + i), f.cur(SIG_HASH + i)));
    a.flush(C);
}
// the light part of schnorr::evaluate_constraints: limb reconstruction of h and the hash copy/injection constraints.
// IN(i): message element injected into the hash at this row.
template <class IN, class CB>
CSG_HD void schnorr_light(const Frame &f, CB &C, fe doubling, fe addition, const fe (&digest)[4], fe copy_hash, IN in) {
    // the four limbs of h are rebuilt from its bits while its scalar multiplication runs (src/schnorr/air.rs:454-486)
    const fe hbit_next = f.next(LIMBS);
    FlagAcc hold(addition);
    CSG_REST_LOOP
    for (int i = 0; i < 4; i++) {
        const int c = LIMBS + 4 - i;
        fe cv = f.cur(c), nv = f.next(c);
        C.add(c, f63::mul(f63::mul(digest[i], doubling), f63::sub(nv, f63::add(f63::dbl(cv), hbit_next))));
        C.add(c, f63::mul(f63::mul(f_not(digest[i]), doubling), f63::sub(cv, nv)));
        hold.add(C, LIMBS + 1 + i, f63::sub(f.cur(LIMBS + 1 + i), f.next(LIMBS + 1 + i)));
    }
    hold.flush(C);
    C.sync();
    // enforce_hash_copy (src/schnorr/air.rs:309-330)
    FlagAcc a(copy_hash);
    CSG_REST_LOOP
    for (int i = 0; i < HRW; i++) {
        a.add(C, SIG_HASH + i, f63::sub(f.cur(SIG_HASH + i), f.next(SIG_HASH + i)));
        a.add(C, SIG_HASH + HRW + i, f63::sub(f.next(SIG_HASH + HRW + i), in(i)));
    }
    a.flush(C);
}

// ================================================================================================ the six AIRs
// Each AIR's evaluate_transition is split into independent work items whose partial sums add up to T(x) (T is linear in
// the result slots): `heavy` items -- one Rescue residual, one scalar-multiplication bank, or the final point addition,
// each a few thousand modular multiplications -- and a `rest` of cheap linear constraints.  The CUDA driver gives every
// heavy item its own thread (small code, moderate registers, 5-8x more parallelism); eval_transition below runs them
// back to back and is what the host-side tests compare with the reference semantics.
template <int AIR> struct Items;
template <> struct Items<TRANSACTION> { static constexpr int rescue = 5, ecc = 3; };
template <> struct Items<MERKLE_UPDATE> { static constexpr int rescue = 4, ecc = 0; };
template <> struct Items<MERKLE_INIT> { static constexpr int rescue = 4, ecc = 0; };
template <> struct Items<SCHNORR> { static constexpr int rescue = 1, ecc = 3; };
template <> struct Items<RANGE> { static constexpr int rescue = 0, ecc = 0; };
template <> struct Items<RESCUE> { static constexpr int rescue = 1, ecc = 0; };

// ---- Rescue residual number s of the AIR: its state columns, round-constant columns and the result slots of its users
struct RescueItem { int col0, ark0, slot_a, slot_b; bool second; };
CSG_HD constexpr RescueItem rescue_item(int air, int s) {
    // transaction AIR: the four leaf/path hash states serve both the leaf-hash phase (setup flag, slots 0,14,28,42:
    // src/merkle/init/air.rs:171-201) and the authentication paths (hash flag, slots = columns); the fifth state is the
    // Schnorr message hash
    return air == TRANSACTION ? (s < 4 ? RescueItem{15 * s - (s >> 1), TX_ARK, 14 * s, 15 * s - (s >> 1), true} : RescueItem{SIG_HASH, TX_ARK, SIG_HASH, 0, false})
         : air == MERKLE_UPDATE ? RescueItem{15 * s - (s >> 1), 5, 15 * s - (s >> 1), 0, false}
         : air == MERKLE_INIT ? RescueItem{15 * s - (s >> 1), 0, 14 * s, 0, false}
         : air == SCHNORR ? RescueItem{SIG_HASH, APW + 15, SIG_HASH, 0, false}
         : RescueItem{0, 1, 0, 0, false};
}
// compiled number of degree groups among the 14 slots of a user (the declared degrees of src/air.rs:76-108 put the slots of
// the curve registers, their bits and the hash states in up to three groups); the host checks the actual count against it
CSG_HD constexpr int rescue_item_ng(int air, int s, int user) {
    return air != TRANSACTION ? 1 : s == 0 ? 2 : s == 1 ? 3 : s == 2 ? 2 : 1;
}

template <int NG, class CB>
CSG_HD void rescue_flush(CB &C, const RescueTables &R, int use, fe flag, const f63::acc192 &a, const f63::acc192 (&b)[NG]) {
    C.sum.mac(flag, a.reduce());
CSG_PRAGMA(unroll)
    for (int q = 0; q < NG; q++)
        if (q < R.ng[use]) {
            const int g = R.grp[use][q];
            if (CB::split) C.part_mac(0, g, flag, b[q].reduce());
            else C.sum.mac(f63::mul(flag, C.xp[g * C.xp_stride]), b[q].reduce());
        }
}
// rescue_state with the forward MDS product folded into per-proof coefficient tables (rescue_tables.h): the alpha part and
// the beta part of each degree group of each user are accumulated unreduced and multiplied by the user's flag once.
template <int NGA, int NGB, class PV, class CB>
CSG_HD void rescue_state_t(const Frame &f, const PV &pv, CB &C, int col0, int ark0, fe flag_a, int use_a, fe flag_b, int use_b) {
    const RescueTables &R = *C.rt;
    constexpr int NB = NGB > 0 ? NGB : 1;
    f63::acc192 aa, ba[NGA], ab, bb[NB];
CSG_PRAGMA(unroll CSG_RESCUE_FWD_UNROLL)
    for (int j = 0; j < 14; j++) {
        const fe t = rescue::cube(f.cur(col0 + j));
        aa.mac(R.a_fwd[use_a][j], t);
CSG_PRAGMA(unroll)
        for (int q = 0; q < NGA; q++) ba[q].mac(R.b_fwd[use_a][q][j], t);
        if (NGB > 0) {
            ab.mac(R.a_fwd[use_b][j], t);
CSG_PRAGMA(unroll)
            for (int q = 0; q < NGB; q++) bb[q].mac(R.b_fwd[use_b][q][j], t);
        }
    }
    fe tn[14];
CSG_PRAGMA(unroll)
    for (int j = 0; j < 14; j++) tn[j] = f63::sub(f.next(col0 + j), pv(ark0 + 14 + j));
    // two rows of the inverse MDS product per iteration: two independent multiply-add chains in flight
CSG_PRAGMA(unroll 1)
    for (int i = 0; i < 14; i += 2) {
        const uint64_t *inv_mds = CSG_TABLE(CSG_INV_MDS) + i * 14;
        f63::acc128 bwd0, bwd1;
CSG_PRAGMA(unroll)
        for (int j = 0; j < 14; j++) { bwd0.mac(inv_mds[j], tn[j]); bwd1.mac(inv_mds[14 + j], tn[j]); }
        const fe v0 = f63::sub(rescue::cube(bwd0.reduce()), pv(ark0 + i)), v1 = f63::sub(rescue::cube(bwd1.reduce()), pv(ark0 + i + 1));
        aa.mac(R.a_bwd[use_a][i], v0); aa.mac(R.a_bwd[use_a][i + 1], v1);
CSG_PRAGMA(unroll)
        for (int q = 0; q < NGA; q++) { ba[q].mac(R.b_bwd[use_a][q][i], v0); ba[q].mac(R.b_bwd[use_a][q][i + 1], v1); }
        if (NGB > 0) {
            ab.mac(R.a_bwd[use_b][i], v0); ab.mac(R.a_bwd[use_b][i + 1], v1);
CSG_PRAGMA(unroll)
            for (int q = 0; q < NGB; q++) { bb[q].mac(R.b_bwd[use_b][q][i], v0); bb[q].mac(R.b_bwd[use_b][q][i + 1], v1); }
        }
    }
    rescue_flush<NGA>(C, R, use_a, flag_a, aa, ba);
    if (NGB > 0) rescue_flush<NB>(C, R, use_b, flag_b, ab, bb);
}

// The tables of every Rescue user of an AIR from the coefficients of one proof (host; rescue_tables.h).  Throws when the
// slots of a user span more degree groups than its accumulators were compiled for.
inline void fill_rescue_tables(int air, const fe *alpha, const fe *beta, const uint8_t *group, unsigned nconstraints, RescueTables &R) {
    R = RescueTables{};
    const int items = air == TRANSACTION ? 5 : air == MERKLE_UPDATE || air == MERKLE_INIT ? 4 : air == SCHNORR || air == RESCUE ? 1 : 0;
    const uint64_t *mds = CSG_MDS_M;
    for (int s = 0; s < items; s++) {
        const RescueItem it = rescue_item(air, s);
        for (int user = 0; user < (it.second ? 2 : 1); user++) {
            const int slot = user ? it.slot_b : it.slot_a, use = 2 * s + user;
            if (use >= RT_MAX_USES || slot + 14 > (int)nconstraints) throw std::runtime_error("Rescue user outside the constraint table");
            unsigned ng = 0;
            int q_of[14];
            for (int i = 0; i < 14; i++) {
                unsigned q = 0;
                while (q < ng && R.grp[use][q] != group[slot + i]) q++;
                if (q == ng) {
                    if (ng == (unsigned)rescue_item_ng(air, s, user)) throw std::runtime_error("more degree groups among the slots of a Rescue user than compiled for");
                    R.grp[use][ng++] = group[slot + i];
                }
                q_of[i] = (int)q;
            }
            R.ng[use] = (unsigned char)ng;
            for (int i = 0; i < 14; i++) { R.a_bwd[use][i] = alpha[slot + i]; R.b_bwd[use][q_of[i]][i] = beta[slot + i]; }
            for (int j = 0; j < 14; j++) {
                fe a = 0, b[RT_MAX_GROUPS] = {0, 0, 0};
                for (int i = 0; i < 14; i++) {
                    a = f63::add(a, f63::mul(alpha[slot + i], mds[i * 14 + j]));
                    b[q_of[i]] = f63::add(b[q_of[i]], f63::mul(beta[slot + i], mds[i * 14 + j]));
                }
                R.a_fwd[use][j] = f63::neg(a);
                for (unsigned q = 0; q < ng; q++) R.b_fwd[use][q][j] = f63::neg(b[q]);
            }
        }
    }
}

template <int AIR, class PV, class CB>
CSG_HD void eval_rescue_item(int s, const Frame &f, const PV &pv, CB &C) {
    const RescueItem it = rescue_item(AIR, s);
    fe flag_a = f63::ONE, flag_b = 0;
    if (AIR == TRANSACTION) { flag_a = s < 4 ? pv(TX_SETUP) : pv(TX_SCHNORR_HASH); flag_b = s < 4 ? pv(TX_HASH) : 0; }
    else if (AIR == MERKLE_UPDATE) flag_a = pv(4);
    else if (AIR == SCHNORR) flag_a = pv(APW + 7);
    else if (AIR == RESCUE) flag_a = pv(0);
#if defined(__CUDA_ARCH__)
    constexpr bool tables = true;    // base field on the device: always through the tables
#else
    const bool tables = C.rt != nullptr;   // host (verifier, tests): the direct form unless tables are supplied
#endif
    if constexpr (CB::D == 1) if (tables) {
        if (AIR == TRANSACTION) {
            // item 0: both users own slots 0..13 -- one user under the sum of the flags
            if (s == 0) rescue_state_t<rescue_item_ng(TRANSACTION, 0, 0), 0>(f, pv, C, it.col0, it.ark0, f63::add(flag_a, flag_b), 0, 0, 0);
            else if (s == 1) rescue_state_t<rescue_item_ng(TRANSACTION, 1, 0), rescue_item_ng(TRANSACTION, 1, 1)>(f, pv, C, it.col0, it.ark0, flag_a, 2, flag_b, 3);
            else if (s == 2) rescue_state_t<rescue_item_ng(TRANSACTION, 2, 0), rescue_item_ng(TRANSACTION, 2, 1)>(f, pv, C, it.col0, it.ark0, flag_a, 4, flag_b, 5);
            else if (s == 3) rescue_state_t<rescue_item_ng(TRANSACTION, 3, 0), rescue_item_ng(TRANSACTION, 3, 1)>(f, pv, C, it.col0, it.ark0, flag_a, 6, flag_b, 7);
            else rescue_state_t<rescue_item_ng(TRANSACTION, 4, 0), 0>(f, pv, C, it.col0, it.ark0, flag_a, 8, 0, 0);
        } else rescue_state_t<1, 0>(f, pv, C, it.col0, it.ark0, flag_a, 2 * s, 0, 0);
        return;
    }
    rescue_state(f, pv, C, it.col0, it.ark0, flag_a, it.slot_a, it.second, flag_b, it.slot_b);
}
// ---- curve items: bank 0 = S (generator), bank 1 = h.P (public key); then the final addition
template <int AIR, class PV, class CB>
CSG_HD void eval_ecc_bank(int bank, const Frame &f, const PV &pv, CB &C) {
    if (AIR != TRANSACTION && AIR != SCHNORR) return;
    const fe scalar_mult = pv(AIR == TRANSACTION ? TX_SCALAR_MULT : 1);
    const fe doubling = pv(AIR == TRANSACTION ? TX_DOUBLING : 2), addition = f63::mul(f_not(doubling), scalar_mult);
    fe q[12];
    const uint64_t *gen = CSG_TABLE(CSG_GENERATOR);
    for (int j = 0; j < 12; j++) q[j] = bank == 0 ? gen[j] : (AIR == TRANSACTION ? f.next(SENDER_KEY + j) : pv(7 + j));
    scalar_mult_bank(f, C, bank * (PPW + 1), q, doubling, addition);
}
template <int AIR, class PV, class CB>
CSG_HD void eval_ecc_bank_outputs(int bank, int formula, const Frame &f, const PV &pv, CB &C) {
    if (AIR != TRANSACTION && AIR != SCHNORR) return;
    fe q[12];
    const uint64_t *gen = CSG_TABLE(CSG_GENERATOR);
    for (int j = 0; j < 12; j++) q[j] = bank == 0 ? gen[j] : (AIR == TRANSACTION ? f.next(SENDER_KEY + j) : pv(7 + j));
    scalar_mult_bank_outputs(f, C, bank * (PPW + 1), q, formula);
}
template <int AIR, class PV, class CB>
CSG_HD fe eval_ecc_bank_merge(int bank, const Frame &f, const PV &pv, const CB &C, int comp, fe Cd, fe Cm) {
    if (AIR != TRANSACTION && AIR != SCHNORR) return 0;
    const fe scalar_mult = pv(AIR == TRANSACTION ? TX_SCALAR_MULT : 1);
    const fe doubling = pv(AIR == TRANSACTION ? TX_DOUBLING : 2), addition = f63::mul(f_not(doubling), scalar_mult);
    return scalar_mult_bank_merge(f, C, comp, bank * (PPW + 1), doubling, addition, Cd, Cm);
}
template <int AIR, class PV, class CB>
CSG_HD void eval_ecc_final(const Frame &f, const PV &pv, CB &C) {
    if (AIR != TRANSACTION && AIR != SCHNORR) return;
    const fe mask = pv(AIR == TRANSACTION ? TX_SCHNORR : 0), scalar_mult = pv(AIR == TRANSACTION ? TX_SCALAR_MULT : 1);
    schnorr_final_addition(f, C, f63::mul(f_not(scalar_mult), mask));
}

template <class PV, class CB>
CSG_HD void rest_transaction(const Frame &f, const PV &pv, CB &C) {
    const fe setup = pv(TX_SETUP), tx_hash = pv(TX_MERKLE), hash_input = pv(TX_HASH_INPUT), finish = pv(TX_FINISH), hashf = pv(TX_HASH);
    const fe schnorr_mask = pv(TX_SCHNORR), scalar_mult = pv(TX_SCALAR_MULT), doubling = pv(TX_DOUBLING), schnorr_hash = pv(TX_SCHNORR_HASH);
    const fe copy_hash = f63::mul(f_not(schnorr_hash), schnorr_mask);
    const fe addition = f63::mul(f_not(doubling), scalar_mult);
    {   // setup row of a transaction: leaf consistency and copies into the carried registers (src/air.rs:405-504)
        FlagAcc a(setup);
        value_block(f, C, a);
        C.sync();
        CSG_REST_LOOP
        for (int o = 0; o < APW; o++) {
            a.add(C, SENDER_KEY_RES + o, f63::sub(f.next(SENDER_KEY + o), f.cur(SENDER_INITIAL + o)));
            a.add(C, RECEIVER_KEY_RES + o, f63::sub(f.next(RECEIVER_KEY + o), f.cur(RECEIVER_INITIAL + o)));
        }
        a.add(C, DELTA_COPY_RES, f63::sub(f.next(DELTA_COPY), f63::sub(f.cur(SENDER_INITIAL + APW), f.cur(SENDER_UPDATED + APW))));
        a.add(C, SIGMA_COPY_RES, f63::sub(f.next(SIGMA_COPY), f.cur(SENDER_UPDATED + APW)));
        a.add(C, NONCE_COPY_RES, f63::sub(f.next(NONCE_COPY), f.cur(SENDER_INITIAL + APW + 1)));
        a.flush(C);
    }
    C.sync();
    {   // carried registers stay put afterwards (src/air.rs:506-529); note the overlapping slot ranges are the reference's
        FlagAcc a(pv(TX_VALUE_COPY));
        CSG_REST_LOOP
        for (int o = 0; o < APW; o++) {
            a.add(C, SENDER_KEY_RES + o, f63::sub(f.next(SENDER_KEY + o), f.cur(SENDER_KEY + o)));
            a.add(C, RECEIVER_KEY_RES + o, f63::sub(f.next(RECEIVER_KEY + o), f.cur(RECEIVER_KEY + o)));
        }
        a.add(C, DELTA_COPY_RES, f63::sub(f.next(DELTA_COPY), f.cur(DELTA_COPY)));
        a.add(C, SIGMA_COPY_RES, f63::sub(f.next(SIGMA_COPY), f.cur(SIGMA_COPY)));
        a.add(C, NONCE_COPY_RES, f63::sub(f.next(NONCE_COPY), f.cur(NONCE_COPY)));
        a.flush(C);
    }
    C.sync();
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int path = 0; path < 2; path++) { merkle_auth_path(f, C, path == 0 ? SENDER_INITIAL : RECEIVER_INITIAL, tx_hash, hash_input, hashf); C.sync(); }
    merkle_roots(f, C, finish);
    C.sync();

    // message elements entering the Schnorr hash come from the carried key/delta/nonce registers (src/air.rs:542-565)
    const fe k0 = pv(TX_INTERNAL), k1 = pv(TX_INTERNAL + 1), k2 = pv(TX_INTERNAL + 2), k3 = pv(TX_INTERNAL + 3);
    auto cell = [&](int idx) -> fe {
        if (idx < 2 * APW) return f.next(SENDER_KEY + idx);   // sender key then receiver key are adjacent columns
        if (idx == 2 * APW) return f.next(DELTA_COPY);
        if (idx == 2 * APW + 1) return f.next(NONCE_COPY);
        return 0;
    };
    auto in = [&](int i) -> fe {
        f63::acc128 t;
        t.mac(k0, cell(i)); t.mac(k1, cell(HRW + i)); t.mac(k2, cell(2 * HRW + i)); t.mac(k3, cell(3 * HRW + i));
        return t.reduce();
    };
    const fe digest[4] = {pv(TX_DIGEST), pv(TX_DIGEST + 1), pv(TX_DIGEST + 2), pv(TX_DIGEST + 3)};
    schnorr_light(f, C, doubling, addition, digest, copy_hash, in);
    C.sync();

    {   // range proofs of delta and sigma (src/air.rs:582-609); the sigma finish check compares the delta registers, as the reference does
        FlagAcc a(pv(TX_RANGE_STEP));
        fe db = f.next(DELTA_BIT), sb = f.next(SIGMA_BIT);
        a.add(C, DELTA_ACC, f63::sub(f.next(DELTA_ACC), f63::add(f63::dbl(f.cur(DELTA_ACC)), db)));
        a.add(C, DELTA_BIT, f_bin(db));
        a.add(C, SIGMA_ACC, f63::sub(f.next(SIGMA_ACC), f63::add(f63::dbl(f.cur(SIGMA_ACC)), sb)));
        a.add(C, SIGMA_BIT, f_bin(sb));
        a.flush(C);
        FlagAcc b(pv(TX_RANGE_FINISH));
        fe v = f63::sub(f.next(DELTA_ACC), f.next(DELTA_COPY));
        b.add(C, DELTA_RANGE_RES, v);
        b.add(C, SIGMA_RANGE_RES, v);
        b.flush(C);
    }
}

// ---- everything that is not a heavy item
template <int AIR, class PV, class CB>
CSG_HD void eval_rest(const Frame &f, const PV &pv, CB &C) {
    if (AIR == TRANSACTION) rest_transaction(f, pv, C);
    else if (AIR == MERKLE_UPDATE) {   // periodic = setup, tx_hash, hash_input, finish, hash, ark[28] (src/merkle/update/air.rs:73-156, 182-212)
        FlagAcc a(pv(0));
        value_block(f, C, a);
        a.flush(C);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int path = 0; path < 2; path++) merkle_auth_path(f, C, path == 0 ? SENDER_INITIAL : RECEIVER_INITIAL, pv(1), pv(2), pv(4));
        merkle_roots(f, C, pv(3));
    } else if (AIR == SCHNORR) {       // periodic = global, scalar_mult, doubling, digest[4], pkey[12], hash flag, chunk[7], ark[28] (src/schnorr/air.rs:75-113)
        const fe global = pv(0), scalar_mult = pv(1), doubling = pv(2), hash_flag = pv(APW + 7);
        const fe digest[4] = {pv(3), pv(4), pv(5), pv(6)};
        schnorr_light(f, C, doubling, f63::mul(f_not(doubling), scalar_mult), digest, f63::mul(f_not(hash_flag), global), [&](int i) { return pv(APW + 8 + i); });
    } else if (AIR == RANGE) {         // src/range/air.rs:69-105: column 0 = bit, column 1 = accumulator
        fe b = f.next(0);
        C.add(1, f63::sub(f.next(1), f63::add(f63::dbl(f.cur(1)), b)));
        C.add(0, f_bin(b));
    } else if (AIR == RESCUE) {        // benches/rescue.rs:205-222: periodic = cycle mask, ark[28]
        FlagAcc a(f_not(pv(0)));
        for (int i = 0; i < HRW; i++) {
            a.add(C, i, f63::sub(f.cur(i), f.next(i)));
            a.add(C, HRW + i, f.next(HRW + i));
        }
        a.flush(C);
    }
    // MERKLE_INIT (src/merkle/init/air.rs:76-90) is its four Rescue rounds only
}

// the whole of Air::evaluate_transition, merged: what one row contributes to T(x)
template <int AIR, class PV, class CB>
CSG_HD void eval_transition(const Frame &f, const PV &pv, CB &C) {
    for (int s = 0; s < Items<AIR>::rescue; s++) eval_rescue_item<AIR>(s, f, pv, C);
    if (Items<AIR>::ecc) { eval_ecc_bank<AIR>(0, f, pv, C); eval_ecc_bank<AIR>(1, f, pv, C); eval_ecc_final<AIR>(f, pv, C); }
    eval_rest<AIR>(f, pv, C);
}

}  // namespace airs
