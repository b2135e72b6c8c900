/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Rescue-XLIX over f63 as the reference defines it: /root/reference/src/utils/rescue.rs
 *   width 14, rate 7, 7 rounds, alpha 3 (rescue.rs:25-37, 381-383); cycle of 8 rows, 8th ARK row is zero (rescue.rs:995)
 */
#ifndef ORACLE_RESCUE_H
#define ORACLE_RESCUE_H
#include "f63.h"

#define RESCUE_STATE_WIDTH 14
#define RESCUE_RATE_WIDTH 7
#define RESCUE_NUM_ROUNDS 7
#define RESCUE_CYCLE 8

void rescue_init_tables(void);                         /* converts the constant tables to Montgomery form (idempotent) */
const fe *rescue_ark(unsigned row);                    /* 28 values, Montgomery */
void rescue_apply_round(fe *state, size_t step);       /* rescue.rs:246-263 */
void rescue_apply_permutation(fe *state);              /* rescue.rs:239-243 */
void rescue_digest(const fe *data, size_t n, fe out[7]); /* rescue.rs:108-130 */
void rescue_merge(const fe a[7], const fe b[7], fe out[7]); /* rescue.rs:143-152 */
/* rescue.rs:269-300 ; result[i] += flag * (inv-half(next) - fwd-half(cur)) */
void rescue_enforce_round(fe *result, const fe *cur, const fe *next, const fe *ark, fe flag);
#endif
