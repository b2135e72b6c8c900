/* ORACLE -- TEST INFRASTRUCTURE ONLY.  The six AIRs of the reference, restated (see air.h for the file map). */
#include "air.h"
#include "ecc.h"
#include "rescue.h"
#include <stdlib.h>
#include <string.h>

/* ---- layout constants: src/merkle/constants.rs:27-56, src/constants.rs:35-99, src/schnorr/constants.rs ---- */
enum {
    HSW = 14, HRW = 7, APW = 12, PPW = 18, PCW = 6,
    SENDER_INITIAL_POS = 0, SENDER_BIT_POS = 14, SENDER_UPDATED_POS = 15, RECEIVER_INITIAL_POS = 29,
    RECEIVER_BIT_POS = 43, RECEIVER_UPDATED_POS = 44, PREV_TREE_ROOT_POS = 58, MERKLE_WIDTH = 65,
    SENDER_INITIAL_RES = 0, RECEIVER_INITIAL_RES = 29, PREV_TREE_ROOT_RES = 58, VALUE_CONSTRAINT_RES = 65,
    BALANCE_CONSTRAINT_RES = 90, NONCE_UPDATE_CONSTRAINT_RES = 91, INT_ROOT_EQUALITY_RES = 92, PREV_TREE_MATCH_RES = 99,
    MERKLE_NUM_CONSTRAINTS = 106, MERKLE_TREE_DEPTH = 15, TRANSACTION_HASH_LENGTH = 8 * 15 + 7, MERKLE_CYCLE = 512,
    SENDER_KEY_POINT_POS = 65, RECEIVER_KEY_POINT_POS = 77, DELTA_COPY_POS = 89, SIGMA_COPY_POS = 90, NONCE_COPY_POS = 91,
    TX_WIDTH = 94, SENDER_KEY_POINT_RES = 101, RECEIVER_KEY_POINT_RES = 103, DELTA_COPY_RES = 105, SIGMA_COPY_RES = 106,
    NONCE_COPY_RES = 107, DELTA_RANGE_RES = 108, SIGMA_RANGE_RES = 109, TX_NUM_CONSTRAINTS = 115,
    SCHNORR_WIDTH = 56, DELTA_BIT_POS = 56, DELTA_ACCUMULATE_POS = 57, SIGMA_BIT_POS = 92, SIGMA_ACCUMULATE_POS = 93,
    TX_CYCLE = 1024, SIG_CYCLE = 512, SCALAR_MUL_LENGTH = 510, NUM_HASH_ITER = 5, RANGE_LOG = 64,
    /* periodic column indices of the transaction AIR: src/constants.rs:85-116 */
    SETUP_MASK = 0, MERKLE_MASK = 1, HASH_INPUT_MASK = 2, FINISH_MASK = 3, HASH_MASK = 4, SCHNORR_MASK = 5,
    SCALAR_MULT_MASK = 6, DOUBLING_MASK = 7, SCHNORR_DIGEST_MASK = 8, SCHNORR_HASH_MASK = 12, HASH_INTERNAL_INPUT_MASKS = 13,
    RANGE_STEP_MASK = 17, RANGE_FINISH_MASK = 18, VALUE_COPY_MASK = 19, ARK_INDEX = 20
};

static inline void agg(fe *r, size_t i, fe flag, fe v) { r[i] = fe_add(r[i], fe_mul(flag, v)); } /* src/utils/mod.rs:58-62 */
static inline fe f_not(fe a) { return fe_sub(FE_ONE, a); }
static inline fe f_bin(fe a) { return fe_sub(fe_sqr(a), a); }

/* src/utils/field.rs:31-70 */
static void enforce_double_and_add(fe *res, const fe *cur, const fe *next, size_t vpos, size_t bpos, fe flag, int with_bit) {
    fe s1 = fe_add(fe_dbl(cur[vpos]), next[bpos]);
    agg(res, vpos, flag, fe_sub(next[vpos], s1));
    if (with_bit) agg(res, bpos, flag, f_bin(next[bpos]));
}

/* ------------------------------------------------------------- merkle::init (src/merkle/init/air.rs:159-202) */
static void merkle_init_constraints(fe *res, const fe *cur, const fe *next, const fe *ark, fe flag) {
    rescue_enforce_round(res + SENDER_INITIAL_POS, cur + SENDER_INITIAL_POS, next + SENDER_INITIAL_POS, ark, flag);
    rescue_enforce_round(res + SENDER_UPDATED_POS - 1, cur + SENDER_UPDATED_POS, next + SENDER_UPDATED_POS, ark, flag);
    rescue_enforce_round(res + RECEIVER_INITIAL_POS - 1, cur + RECEIVER_INITIAL_POS, next + RECEIVER_INITIAL_POS, ark, flag);
    rescue_enforce_round(res + RECEIVER_UPDATED_POS - 2, cur + RECEIVER_UPDATED_POS, next + RECEIVER_UPDATED_POS, ark, flag);
}

/* ------------------------------------------------------------- merkle::update (src/merkle/update/air.rs:291-369) */
static void merkle_update_auth(fe *res, const fe *cur, const fe *next, const fe *ark, fe tx_hash_flag, fe hash_input_flag, fe hash_flag) {
    fe copy_flag = fe_mul(tx_hash_flag, f_not(fe_add(hash_flag, hash_input_flag)));
    fe init_flag = fe_mul(tx_hash_flag, hash_input_flag);
    fe bit = next[HSW], nbit = f_not(bit);
    agg(res, HSW, tx_hash_flag, f_bin(bit));
    const size_t offs[2] = {0, HSW + 1};
    for (int k = 0; k < 2; k++) {
        size_t o = offs[k];
        rescue_enforce_round(res + o, cur + o, next + o, ark, hash_flag);
        for (size_t i = 0; i < HRW; i++) {
            agg(res, o + i, copy_flag, fe_sub(cur[o + i], next[o + i]));
            agg(res, o + i, init_flag, fe_mul(nbit, fe_sub(cur[o + i], next[o + i])));
            agg(res, o + HRW + i, init_flag, fe_mul(bit, fe_sub(cur[o + i], next[o + HRW + i])));
        }
    }
    for (size_t i = 0; i < HRW; i++) agg(res, i, init_flag, fe_mul(bit, fe_sub(next[HSW + 1 + i], next[i])));
    for (size_t i = HRW; i < HSW; i++) agg(res, i, init_flag, fe_mul(nbit, fe_sub(next[HSW + 1 + i], next[i])));
}
/* src/merkle/update/air.rs:215-289 */
static void merkle_update_constraints(fe *res, const fe *cur, const fe *next, const fe *ark, fe tx_hash_flag, fe hash_input_flag,
                                      fe hash_flag, fe finish_flag) {
    fe not_finish = f_not(finish_flag);
    merkle_update_auth(res + SENDER_INITIAL_RES, cur + SENDER_INITIAL_POS, next + SENDER_INITIAL_POS, ark, tx_hash_flag, hash_input_flag, hash_flag);
    merkle_update_auth(res + RECEIVER_INITIAL_RES, cur + RECEIVER_INITIAL_POS, next + RECEIVER_INITIAL_POS, ark, tx_hash_flag, hash_input_flag, hash_flag);
    for (size_t i = 0; i < HRW; i++) {
        agg(res, PREV_TREE_ROOT_RES + i, not_finish, fe_sub(next[PREV_TREE_ROOT_POS + i], cur[PREV_TREE_ROOT_POS + i]));
        agg(res, PREV_TREE_ROOT_RES + i, finish_flag, fe_sub(next[PREV_TREE_ROOT_POS + i], next[RECEIVER_UPDATED_POS + i]));
    }
    for (size_t i = 0; i < HRW; i++)
        agg(res, INT_ROOT_EQUALITY_RES + i, finish_flag, fe_sub(cur[SENDER_UPDATED_POS + i], cur[RECEIVER_INITIAL_POS + i]));
    for (size_t i = 0; i < HRW; i++)
        agg(res, PREV_TREE_MATCH_RES + i, finish_flag, fe_sub(next[SENDER_INITIAL_POS + i], cur[PREV_TREE_ROOT_POS + i]));
}
/* value / balance / nonce block shared by MerkleAir::evaluate_transition (update/air.rs:96-144) and the tx AIR (src/air.rs:405-453) */
static void value_constraints(fe *res, const fe *cur, fe setup_flag) {
    for (size_t i = 0; i < APW; i++) {
        agg(res, VALUE_CONSTRAINT_RES + i, setup_flag, fe_sub(cur[SENDER_INITIAL_POS + i], cur[SENDER_UPDATED_POS + i]));
        agg(res, VALUE_CONSTRAINT_RES + APW + i, setup_flag, fe_sub(cur[RECEIVER_INITIAL_POS + i], cur[RECEIVER_UPDATED_POS + i]));
    }
    agg(res, VALUE_CONSTRAINT_RES + 2 * APW, setup_flag, fe_sub(cur[RECEIVER_INITIAL_POS + APW + 1], cur[RECEIVER_UPDATED_POS + APW + 1]));
    agg(res, BALANCE_CONSTRAINT_RES, setup_flag,
        fe_sub(fe_sub(cur[SENDER_INITIAL_POS + APW], cur[SENDER_UPDATED_POS + APW]), fe_sub(cur[RECEIVER_UPDATED_POS + APW], cur[RECEIVER_INITIAL_POS + APW])));
    agg(res, NONCE_UPDATE_CONSTRAINT_RES, setup_flag, fe_sub(cur[SENDER_UPDATED_POS + APW + 1], fe_add(cur[SENDER_INITIAL_POS + APW + 1], FE_ONE)));
}

/* ------------------------------------------------------------- schnorr (src/schnorr/air.rs:394-531, 309-330) */
static void schnorr_constraints(fe *res, const fe *cur, const fe *next, const fe *ark, fe doubling_flag, fe addition_flag,
                                const fe *digest_flags, const fe *pkey, fe final_add_flag, fe hash_flag, fe copy_hash_flag,
                                const fe *internal_inputs) {
    ecc_enforce_doubling(res, cur, next, doubling_flag);
    ecc_enforce_addition_mixed(res, cur, next, ecc_generator(), addition_flag);
    ecc_enforce_doubling(res + PPW + 1, cur + PPW + 1, next + PPW + 1, doubling_flag);
    ecc_enforce_addition_mixed(res + PPW + 1, cur + PPW + 1, next + PPW + 1, pkey, addition_flag);
    const size_t L = 2 * PPW + 1; /* 37: h-bit, then 4 limb accumulators */
    for (size_t i = 0; i < 4; i++)
        enforce_double_and_add(res + L, cur + L, next + L, 4 - i, 0, fe_mul(digest_flags[i], doubling_flag), 0);
    for (size_t i = 0; i < 4; i++) agg(res, L + 1 + i, addition_flag, fe_sub(cur[L + 1 + i], next[L + 1 + i]));
    for (size_t i = 0; i < 4; i++)
        agg(res, L + 4 - i, fe_mul(f_not(digest_flags[i]), doubling_flag), fe_sub(cur[L + 4 - i], next[L + 4 - i]));
    const size_t H = 2 * PPW + 6; /* 42: Rescue state */
    rescue_enforce_round(res + H, cur + H, next + H, ark, hash_flag);
    for (size_t i = 0; i < HRW; i++) agg(res, H + i, copy_hash_flag, fe_sub(cur[H + i], next[H + i]));
    for (size_t i = 0; i < HRW; i++) agg(res, H + HRW + i, copy_hash_flag, fe_sub(next[H + HRW + i], internal_inputs[i]));
    ecc_enforce_addition_reduce_x(res, cur, next, cur + PPW + 1, final_add_flag);
    for (size_t i = 0; i < 4; i++) agg(res, L + 1 + i, final_add_flag, fe_sub(cur[L + 1 + i], cur[H + i]));
}

/* ------------------------------------------------------------- evaluate_transition of each AIR */
/* src/air.rs:114-173 + 383-610 */
static void eval_transaction(const air_t *a, const fe *cur, const fe *next, const fe *pv, fe *res) {
    (void)a;
    fe setup = pv[SETUP_MASK], tx_hash = pv[MERKLE_MASK], hash_input = pv[HASH_INPUT_MASK], finish = pv[FINISH_MASK], hashf = pv[HASH_MASK];
    fe schnorr_mask = pv[SCHNORR_MASK], scalar_mult = pv[SCALAR_MULT_MASK], doubling = pv[DOUBLING_MASK];
    const fe *digest_flags = pv + SCHNORR_DIGEST_MASK;
    fe schnorr_hash = pv[SCHNORR_HASH_MASK];
    const fe *internal_flags = pv + HASH_INTERNAL_INPUT_MASKS;
    fe range_flag = pv[RANGE_STEP_MASK], range_finish = pv[RANGE_FINISH_MASK], copy_values = pv[VALUE_COPY_MASK];
    const fe *ark = pv + ARK_INDEX;
    fe copy_hash = fe_mul(f_not(schnorr_hash), schnorr_mask);
    fe final_add = fe_mul(f_not(scalar_mult), schnorr_mask);
    fe addition = fe_mul(f_not(doubling), scalar_mult);

    merkle_init_constraints(res, cur, next, ark, setup);
    value_constraints(res, cur, setup);
    /* key / delta / sigma / nonce copies at the start of the transaction: src/air.rs:455-504 */
    for (size_t o = 0; o < APW; o++) {
        agg(res, SENDER_KEY_POINT_RES + o, setup, fe_sub(next[SENDER_KEY_POINT_POS + o], cur[SENDER_INITIAL_POS + o]));
        agg(res, RECEIVER_KEY_POINT_RES + o, setup, fe_sub(next[RECEIVER_KEY_POINT_POS + o], cur[RECEIVER_INITIAL_POS + o]));
    }
    agg(res, DELTA_COPY_RES, setup, fe_sub(next[DELTA_COPY_POS], fe_sub(cur[SENDER_INITIAL_POS + APW], cur[SENDER_UPDATED_POS + APW])));
    agg(res, SIGMA_COPY_RES, setup, fe_sub(next[SIGMA_COPY_POS], cur[SENDER_UPDATED_POS + APW]));
    agg(res, NONCE_COPY_RES, setup, fe_sub(next[NONCE_COPY_POS], cur[SENDER_INITIAL_POS + APW + 1]));
    /* ... and for the remainder of the transaction: src/air.rs:506-529 */
    for (size_t o = 0; o < APW; o++) {
        agg(res, SENDER_KEY_POINT_RES + o, copy_values, fe_sub(next[SENDER_KEY_POINT_POS + o], cur[SENDER_KEY_POINT_POS + o]));
        agg(res, RECEIVER_KEY_POINT_RES + o, copy_values, fe_sub(next[RECEIVER_KEY_POINT_POS + o], cur[RECEIVER_KEY_POINT_POS + o]));
    }
    agg(res, DELTA_COPY_RES, copy_values, fe_sub(next[DELTA_COPY_POS], cur[DELTA_COPY_POS]));
    agg(res, SIGMA_COPY_RES, copy_values, fe_sub(next[SIGMA_COPY_POS], cur[SIGMA_COPY_POS]));
    agg(res, NONCE_COPY_RES, copy_values, fe_sub(next[NONCE_COPY_POS], cur[NONCE_COPY_POS]));

    merkle_update_constraints(res, cur, next, ark, tx_hash, hash_input, hashf, finish);

    /* message chunks injected into the Schnorr hash: src/air.rs:542-565 */
    fe inputs[HRW] = {0};
    for (size_t k = 0; k < NUM_HASH_ITER - 1; k++)
        for (size_t i = 0; i < HRW; i++) {
            size_t idx = k * HRW + i;
            fe cell = 0;
            if (idx < APW) cell = next[SENDER_KEY_POINT_POS + idx];
            else if (idx < 2 * APW) cell = next[RECEIVER_KEY_POINT_POS + idx - APW];
            else if (idx == 2 * APW) cell = next[DELTA_COPY_POS];
            else if (idx == 2 * APW + 1) cell = next[NONCE_COPY_POS];
            inputs[i] = fe_add(inputs[i], fe_mul(internal_flags[k], cell));
        }
    schnorr_constraints(res, cur, next, ark, doubling, addition, digest_flags, next + SENDER_KEY_POINT_POS, final_add, schnorr_hash, copy_hash, inputs);

    enforce_double_and_add(res, cur, next, DELTA_ACCUMULATE_POS, DELTA_BIT_POS, range_flag, 1);
    enforce_double_and_add(res, cur, next, SIGMA_ACCUMULATE_POS, SIGMA_BIT_POS, range_flag, 1);
    /* src/air.rs:600-609 -- both finish constraints compare the DELTA registers (reference quirk, kept) */
    agg(res, DELTA_RANGE_RES, range_finish, fe_sub(next[DELTA_ACCUMULATE_POS], next[DELTA_COPY_POS]));
    agg(res, SIGMA_RANGE_RES, range_finish, fe_sub(next[DELTA_ACCUMULATE_POS], next[DELTA_COPY_POS]));
}
/* src/merkle/update/air.rs:73-156 */
static void eval_merkle_update(const air_t *a, const fe *cur, const fe *next, const fe *pv, fe *res) {
    (void)a;
    value_constraints(res, cur, pv[0]);
    merkle_update_constraints(res, cur, next, pv + 5, pv[1], pv[2], pv[4], pv[3]);
}
/* src/merkle/init/air.rs:76-90 */
static void eval_merkle_init(const air_t *a, const fe *cur, const fe *next, const fe *pv, fe *res) {
    (void)a;
    merkle_init_constraints(res, cur, next, pv, FE_ONE);
}
/* src/schnorr/air.rs:75-113 */
static void eval_schnorr(const air_t *a, const fe *cur, const fe *next, const fe *pv, fe *res) {
    (void)a;
    fe global = pv[0], scalar_mult = pv[1], doubling = pv[2], hash_flag = pv[APW + 7];
    fe copy_hash = fe_mul(f_not(hash_flag), global), final_add = fe_mul(f_not(scalar_mult), global);
    fe addition = fe_mul(f_not(doubling), scalar_mult);
    schnorr_constraints(res, cur, next, pv + APW + 15, doubling, addition, pv + 3, pv + 7, final_add, hash_flag, copy_hash, pv + APW + 8);
}
/* src/range/air.rs:69-105 */
static void eval_range(const air_t *a, const fe *cur, const fe *next, const fe *pv, fe *res) {
    (void)a; (void)pv;
    enforce_double_and_add(res, cur, next, 1, 0, FE_ONE, 1);
}
/* benches/rescue.rs:205-222, 256-268 */
static void eval_rescue(const air_t *a, const fe *cur, const fe *next, const fe *pv, fe *res) {
    (void)a;
    fe hash_flag = pv[0], copy_flag = f_not(pv[0]);
    rescue_enforce_round(res, cur, next, pv + 1, hash_flag);
    for (size_t i = 0; i < HRW; i++) agg(res, i, copy_flag, fe_sub(cur[i], next[i]));
    for (size_t i = 0; i < HRW; i++) agg(res, HRW + i, copy_flag, next[HRW + i]);
}

/* ------------------------------------------------------------- periodic columns */
typedef struct { fe *v; size_t len, cap; } col_t;
static void col_push(col_t *c, fe x) {
    if (c->len == c->cap) { c->cap = c->cap ? c->cap * 2 : 16; c->v = realloc(c->v, c->cap * sizeof(fe)); }
    c->v[c->len++] = x;
}
static void col_pad(col_t *c, size_t len, fe x) { while (c->len < len) col_push(c, x); }   /* periodic_columns.rs:185-214 */
static void col_append(col_t *c, const fe *src, size_t n) { for (size_t i = 0; i < n; i++) col_push(c, src[i]); } /* stitch :54-75 */
/* the 28 round-constant columns of length 8: rescue.rs:303-318 */
static void ark_columns(col_t *cols) {
    for (size_t j = 0; j < 2 * HSW; j++) for (size_t i = 0; i < 8; i++) col_push(&cols[j], rescue_ark(i)[j]);
}
static void set_periodic(air_t *a, col_t *cols, uint32_t n) {
    a->num_periodic = n;
    a->periodic = malloc(n * sizeof(fe *));
    a->periodic_len = malloc(n * sizeof(size_t));
    for (uint32_t i = 0; i < n; i++) { a->periodic[i] = cols[i].v; a->periodic_len[i] = cols[i].len; }
}
/* merkle::update::periodic_columns() masks (update/air.rs:182-212), for an arbitrary tree depth */
static void merkle_masks(col_t *setup, col_t *tx_hash, col_t *hash_input, col_t *finish, col_t *hashm, size_t hash_len, size_t upto) {
    for (size_t i = 0; i < upto; i++) {
        if (setup) col_push(setup, i == 0 ? FE_ONE : FE_ZERO);
        fe th = i < hash_len ? FE_ONE : FE_ZERO;
        col_push(tx_hash, th);
        col_push(finish, i == hash_len - 1 ? FE_ONE : FE_ZERO);
        col_push(hashm, (i % 8) < 7 ? th : FE_ZERO);
    }
    if (hash_input) for (size_t i = 0; i < 8; i++) col_push(hash_input, i == 7 ? FE_ONE : FE_ZERO);
}
/* schnorr::periodic_columns() masks (schnorr/air.rs:334-391): global, scalar_mult, doubling, digest x4, hash_flag */
static void schnorr_masks(col_t *c /* 8 columns */) {
    for (size_t i = 0; i < SIG_CYCLE; i++) {
        col_push(&c[0], i < SCALAR_MUL_LENGTH + 1 ? FE_ONE : FE_ZERO);
        col_push(&c[1], i < SCALAR_MUL_LENGTH ? FE_ONE : FE_ZERO);
        col_push(&c[2], (i < SCALAR_MUL_LENGTH && i % 2 == 0) ? FE_ONE : FE_ZERO);
        col_push(&c[3], i < 126 ? FE_ONE : FE_ZERO);
        col_push(&c[4], (i >= 126 && i < 254) ? FE_ONE : FE_ZERO);
        col_push(&c[5], (i >= 254 && i < 382) ? FE_ONE : FE_ZERO);
        col_push(&c[6], (i >= 382 && i < 510) ? FE_ONE : FE_ZERO);
        col_push(&c[7], (i < 8 * NUM_HASH_ITER && i % 8 < 7) ? FE_ONE : FE_ZERO);
    }
}
/* src/air.rs:194-380 */
static void periodic_transaction(air_t *a) {
    col_t *c = calloc(48, sizeof(col_t));
    ark_columns(c + ARK_INDEX);
    col_pad(&c[SETUP_MASK], 1, FE_ONE);
    col_pad(&c[VALUE_COPY_MASK], 1, FE_ZERO);
    merkle_masks(NULL, &c[MERKLE_MASK], &c[HASH_INPUT_MASK], &c[FINISH_MASK], &c[HASH_MASK], TRANSACTION_HASH_LENGTH, TRANSACTION_HASH_LENGTH);
    const int zpad512[] = {SETUP_MASK, MERKLE_MASK, FINISH_MASK, HASH_MASK, SCHNORR_MASK, SCALAR_MULT_MASK, DOUBLING_MASK, 8, 9, 10, 11,
                           SCHNORR_HASH_MASK, 13, 14, 15, 16, RANGE_STEP_MASK, RANGE_FINISH_MASK};
    for (size_t i = 0; i < sizeof zpad512 / sizeof *zpad512; i++) col_pad(&c[zpad512[i]], MERKLE_CYCLE, FE_ZERO);
    col_pad(&c[VALUE_COPY_MASK], MERKLE_CYCLE, FE_ONE);
    col_t s[8] = {{0}};
    schnorr_masks(s);
    const int smap[8] = {SCHNORR_MASK, SCALAR_MULT_MASK, DOUBLING_MASK, 8, 9, 10, 11, SCHNORR_HASH_MASK};
    for (int i = 0; i < 8; i++) { col_append(&c[smap[i]], s[i].v, s[i].len); free(s[i].v); }
    for (size_t k = 0; k < NUM_HASH_ITER - 1; k++)
        for (size_t i = 0; i < SIG_CYCLE; i++) col_push(&c[HASH_INTERNAL_INPUT_MASKS + k], i == (k + 1) * 8 - 1 ? FE_ONE : FE_ZERO);
    for (size_t i = 0; i < RANGE_LOG; i++) { col_push(&c[RANGE_STEP_MASK], FE_ONE); col_push(&c[RANGE_FINISH_MASK], i == RANGE_LOG - 1 ? FE_ONE : FE_ZERO); }
    col_pad(&c[VALUE_COPY_MASK], MERKLE_CYCLE + RANGE_LOG, FE_ONE);
    for (int i = 0; i < ARK_INDEX; i++) if (i != HASH_INPUT_MASK) col_pad(&c[i], TX_CYCLE, FE_ZERO);
    set_periodic(a, c, 48);
    free(c);
}
static void periodic_merkle_update(air_t *a) {
    col_t *c = calloc(33, sizeof(col_t));
    merkle_masks(&c[0], &c[1], &c[2], &c[3], &c[4], TRANSACTION_HASH_LENGTH, MERKLE_CYCLE);
    ark_columns(c + 5);
    set_periodic(a, c, 33);
    free(c);
}
static void periodic_ark_only(air_t *a, int with_cycle_mask) {
    col_t *c = calloc(29, sizeof(col_t));
    int o = 0;
    if (with_cycle_mask) { for (size_t i = 0; i < 8; i++) col_push(&c[0], i < 7 ? FE_ONE : FE_ZERO); o = 1; } /* benches/rescue.rs:119-128 */
    ark_columns(c + o);
    set_periodic(a, c, 28 + o);
    free(c);
}
/* src/schnorr/air.rs:229-299 */
static void periodic_schnorr(air_t *a) {
    col_t *c = calloc(55, sizeof(col_t));
    col_t s[8] = {{0}};
    schnorr_masks(s);
    for (int i = 0; i < 7; i++) c[i] = s[i];
    c[7 + APW] = s[7];
    size_t n = SIG_CYCLE * a->nsig;
    for (size_t j = 0; j < APW; j++) col_pad(&c[7 + j], n, FE_ZERO);
    for (size_t j = 0; j < HRW; j++) col_pad(&c[8 + APW + j], n, FE_ZERO);
    for (size_t m = 0; m < a->nsig; m++) {
        const uint64_t *msg = a->pub_inputs + m * 38;
        for (size_t i = 0; i < NUM_HASH_ITER - 1; i++)
            for (size_t j = 0; j < HRW; j++) c[8 + APW + j].v[i * 8 + 7 + m * SIG_CYCLE] = fe_from_u64(msg[j + i * HRW]);
        for (size_t i = 0; i < SIG_CYCLE; i++)
            for (size_t j = 0; j < APW; j++) c[7 + j].v[i + m * SIG_CYCLE] = fe_from_u64(msg[j]);
    }
    ark_columns(c + 27);
    set_periodic(a, c, 55);
    free(c);
}

/* ------------------------------------------------------------- degrees */
static air_degree deg(uint32_t base, uint32_t nc, uint32_t c) { air_degree d = {base, nc, {c, c}}; return d; }
/* src/merkle/update/air.rs:371-401 */
static void merkle_update_degrees(air_degree *d, uint32_t cyc) {
    for (int half = 0; half < 2; half++) {
        air_degree *h = d + half * 29;
        for (int i = 0; i < 14; i++) h[i] = deg(3, 1, cyc);
        h[14] = deg(2, 1, cyc);
        for (int i = 15; i < 29; i++) h[i] = deg(3, 1, cyc);
    }
    for (int i = 58; i < MERKLE_NUM_CONSTRAINTS; i++) d[i] = deg(1, 1, cyc);
}
/* src/schnorr/air.rs:533-585 */
static void schnorr_degrees(air_degree *d, size_t num_tx, uint32_t cyc) {
    uint32_t bit_degree = num_tx == 1 ? 3 : 5;
    int k = 0;
    for (int i = 0; i < PCW; i++) d[k++] = deg(5, 2, cyc);
    for (int i = 0; i < APW; i++) d[k++] = deg(4, 2, cyc);
    d[k++] = deg(2, 1, cyc);
    for (int i = 0; i < PPW; i++) d[k++] = deg(bit_degree, 2, cyc);
    d[k++] = deg(2, 1, cyc);
    for (int i = 0; i < 4; i++) d[k++] = deg(1, 2, cyc);
    for (int i = 0; i < HSW; i++) d[k++] = deg(3, 1, cyc);
}

/* ------------------------------------------------------------- constructors */
static void add_assert(air_t *a, uint32_t col, size_t first, size_t stride, const fe *vals, size_t nv) {
    a->assertions = realloc(a->assertions, (a->num_assertions + 1) * sizeof(air_assertion));
    air_assertion *s = &a->assertions[a->num_assertions++];
    s->column = col; s->first_step = first;
    s->stride = stride;
    s->nvalues = nv;
    s->values = malloc(nv * sizeof(fe));
    memcpy(s->values, vals, nv * sizeof(fe));
}
static void add_single(air_t *a, uint32_t col, size_t step, fe v) { add_assert(a, col, step, 0, &v, 1); }
static void add_periodic(air_t *a, uint32_t col, size_t first, size_t stride, fe v) { add_assert(a, col, first, stride, &v, 1); }
/* winterfell Assertion::sequence: a one-value sequence degenerates to a single assertion (stride 0) [RECALLED] */
static void add_sequence(air_t *a, uint32_t col, size_t first, size_t stride, const fe *v, size_t nv) { add_assert(a, col, first, nv == 1 ? 0 : stride, v, nv); }

air_t *air_new(int id, size_t n, const uint64_t *pub, size_t npub) {
    rescue_init_tables(); ecc_init_tables();
    air_t *a = calloc(1, sizeof *a);
    a->id = id; a->trace_len = n;
    a->pub_inputs = malloc((npub ? npub : 1) * sizeof(uint64_t));
    memcpy(a->pub_inputs, pub, npub * sizeof(uint64_t));
    a->num_pub_inputs = npub;
    switch (id) {
    case AIR_TRANSACTION: { /* src/air.rs:76-108, 175-184 */
        a->width = TX_WIDTH; a->num_constraints = TX_NUM_CONSTRAINTS; a->eval = eval_transaction;
        a->degrees = calloc(TX_NUM_CONSTRAINTS, sizeof(air_degree));
        merkle_update_degrees(a->degrees, TX_CYCLE);
        a->degrees[RECEIVER_BIT_POS] = deg(3, 1, TX_CYCLE);
        a->degrees[INT_ROOT_EQUALITY_RES] = deg(2, 1, TX_CYCLE);
        air_degree sd[SCHNORR_WIDTH];
        schnorr_degrees(sd, 2, TX_CYCLE);
        for (int i = 0; i < PPW; i++) { a->degrees[i] = sd[i]; a->degrees[i + PPW + 1] = sd[i + PPW + 1]; }
        for (int i = MERKLE_NUM_CONSTRAINTS; i < TX_NUM_CONSTRAINTS; i++) a->degrees[i] = deg(1, 1, TX_CYCLE);
        periodic_transaction(a);
        add_single(a, PREV_TREE_ROOT_POS, 0, fe_from_u64(pub[0]));
        add_single(a, PREV_TREE_ROOT_POS + 1, 0, fe_from_u64(pub[1]));
        add_single(a, PREV_TREE_ROOT_POS, n - 1, fe_from_u64(pub[7]));
        add_single(a, PREV_TREE_ROOT_POS + 1, n - 1, fe_from_u64(pub[8]));
        break;
    }
    case AIR_MERKLE_UPDATE: { /* update/air.rs:46-56, 158-177 */
        a->width = MERKLE_WIDTH; a->num_constraints = MERKLE_NUM_CONSTRAINTS; a->eval = eval_merkle_update;
        a->degrees = calloc(MERKLE_NUM_CONSTRAINTS, sizeof(air_degree));
        merkle_update_degrees(a->degrees, MERKLE_CYCLE);
        periodic_merkle_update(a);
        for (int i = 0; i < HRW; i++) add_single(a, PREV_TREE_ROOT_POS + i, 0, fe_from_u64(pub[i]));
        for (int i = 0; i < HRW; i++) add_single(a, PREV_TREE_ROOT_POS + i, n - 1, fe_from_u64(pub[7 + i]));
        break;
    }
    case AIR_MERKLE_INIT: { /* init/air.rs:50-150, 204-211 */
        a->width = 58; a->num_constraints = 56; a->eval = eval_merkle_init;
        a->degrees = calloc(56, sizeof(air_degree));
        for (int i = 0; i < 56; i++) a->degrees[i] = deg(3, 0, 0);
        periodic_ark_only(a, 0);
        fe s[14], r[14], delta = fe_from_u64(pub[28]);
        for (int i = 0; i < 14; i++) { s[i] = fe_from_u64(pub[i]); r[i] = fe_from_u64(pub[14 + i]); }
        for (int i = 0; i < APW + 2; i++) add_single(a, SENDER_INITIAL_POS + i, 0, s[i]);
        for (int i = 0; i < APW; i++) add_single(a, SENDER_UPDATED_POS + i, 0, s[i]);
        add_single(a, SENDER_UPDATED_POS + APW, 0, fe_sub(s[APW], delta));
        add_single(a, SENDER_UPDATED_POS + APW + 1, 0, fe_add(s[APW + 1], FE_ONE));
        for (int i = 0; i < APW + 2; i++) add_single(a, RECEIVER_INITIAL_POS + i, 0, r[i]);
        for (int i = 0; i < APW; i++) add_single(a, RECEIVER_UPDATED_POS + i, 0, r[i]);
        add_single(a, RECEIVER_UPDATED_POS + APW, 0, fe_add(r[APW], delta));
        add_single(a, RECEIVER_UPDATED_POS + APW + 1, 0, r[APW + 1]);
        break;
    }
    case AIR_SCHNORR: { /* schnorr/air.rs:50-58, 115-227 */
        a->nsig = npub / 38;
        a->width = SCHNORR_WIDTH; a->num_constraints = SCHNORR_WIDTH; a->eval = eval_schnorr;
        a->degrees = calloc(SCHNORR_WIDTH, sizeof(air_degree));
        schnorr_degrees(a->degrees, a->nsig, SIG_CYCLE);
        periodic_schnorr(a);
        for (int i = 0; i < PPW; i++) add_periodic(a, i, 0, SIG_CYCLE, i == PCW ? FE_ONE : FE_ZERO);
        add_periodic(a, PPW, 0, SIG_CYCLE, FE_ZERO);
        for (int i = 0; i < PPW; i++) add_periodic(a, i + PPW + 1, 0, SIG_CYCLE, i == PCW ? FE_ONE : FE_ZERO);
        for (int i = 0; i < 5; i++) add_periodic(a, i + 2 * PPW + 1, 0, SIG_CYCLE, FE_ZERO);
        fe *rx = malloc(a->nsig * sizeof(fe));
        for (int l = 0; l < PCW; l++) {
            for (size_t m = 0; m < a->nsig; m++) rx[m] = fe_from_u64(pub[m * 38 + 28 + l]);
            add_sequence(a, 2 * PPW + 6 + l, 0, SIG_CYCLE, rx, a->nsig);
        }
        for (int i = 0; i < HRW; i++) add_periodic(a, i + 2 * PPW + PCW + 6, 0, SIG_CYCLE, FE_ZERO);
        for (int l = 0; l < PCW; l++) {
            for (size_t m = 0; m < a->nsig; m++) rx[m] = fe_from_u64(pub[m * 38 + 28 + l]);
            add_sequence(a, l, SCALAR_MUL_LENGTH + 1, SIG_CYCLE, rx, a->nsig);
        }
        free(rx);
        break;
    }
    case AIR_RANGE: { /* range/air.rs:43-105 */
        a->width = 2; a->num_constraints = 2; a->eval = eval_range;
        a->degrees = calloc(2, sizeof(air_degree));
        a->degrees[0] = deg(2, 0, 0); a->degrees[1] = deg(1, 0, 0);
        add_single(a, 1, 0, FE_ZERO);
        add_single(a, 1, n - 1, fe_from_u64(pub[0]));
        break;
    }
    case AIR_RESCUE: { /* benches/rescue.rs:163-254 */
        a->width = 14; a->num_constraints = 14; a->eval = eval_rescue;
        a->degrees = calloc(14, sizeof(air_degree));
        for (int i = 0; i < 14; i++) a->degrees[i] = deg(3, 1, 8);
        periodic_ark_only(a, 1);
        for (int i = 0; i < 7; i++) add_single(a, i, 0, fe_from_u64(pub[i]));
        for (int i = 0; i < 7; i++) add_single(a, i, n - 1, fe_from_u64(pub[7 + i]));
        break;
    }
    default: free(a->pub_inputs); free(a); return NULL;
    }
    return a;
}
void air_free(air_t *a) {
    if (!a) return;
    for (uint32_t i = 0; i < a->num_periodic; i++) free(a->periodic[i]);
    for (uint32_t i = 0; i < a->num_assertions; i++) free(a->assertions[i].values);
    free(a->periodic); free(a->periodic_len); free(a->assertions); free(a->degrees); free(a->pub_inputs); free(a);
}
void air_eval_row(const air_t *a, size_t step, const fe *cur, const fe *next, fe *result) {
    fe pv[64];
    for (uint32_t i = 0; i < a->num_periodic; i++) pv[i] = a->periodic[i][step % a->periodic_len[i]];
    memset(result, 0, a->num_constraints * sizeof(fe));
    a->eval(a, cur, next, pv, result);
}
/* winterfell TransitionConstraintDegree::get_evaluation_degree [RECALLED]:
 * base * (n - 1) + sum over cycles of (n / cycle) * (cycle - 1) */
size_t air_eval_degree(const air_degree *d, size_t n) {
    size_t r = (size_t)d->base * (n - 1);
    for (uint32_t i = 0; i < d->ncycles; i++) r += (n / d->cycles[i]) * (d->cycles[i] - 1);
    return r;
}
/* winterfell AirContext::new [RECALLED]: ce_blowup = max over constraints of next_pow2(base + #cycles), at least 2 */
size_t air_ce_blowup(const air_t *a) {
    size_t m = 2;
    for (uint32_t i = 0; i < a->num_constraints; i++) {
        size_t v = a->degrees[i].base + a->degrees[i].ncycles, p2 = 1;
        while (p2 < v) p2 *= 2;
        if (p2 > m) m = p2;
    }
    return m;
}
