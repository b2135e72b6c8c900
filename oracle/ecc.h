/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Curve arithmetic used by the Schnorr sub-AIR: /root/reference/src/utils/ecc.rs
 *   Fp2 = Fp[u]/(u^2 - 2u - 2)            (ecc.rs:407-446, derived from mul_fp2/square_fp2)
 *   Fp6 = Fp2[v]/(v^3 + v + 1)            (ecc.rs:462-548, derived from mul_fp6/square_fp6)
 *   E: y^2 = x^3 + x + B, B3 = 3B         (ecc.rs:38-45), complete projective formulas of
 *   Renes-Costello-Batina 2015 (Alg. 1 add, Alg. 2 mixed add, Alg. 3 doubling) with a = 1 (ecc.rs:186-404)
 */
#ifndef ORACLE_ECC_H
#define ORACLE_ECC_H
#include "f63.h"

void ecc_init_tables(void);
const fe *ecc_generator(void); /* 12 limbs, Montgomery (ecc.rs:23-36) */
void fp6_mul(fe r[6], const fe a[6], const fe b[6]);
void fp6_sqr(fe r[6], const fe a[6]);
void fp6_inv(fe r[6], const fe a[6]);
/* in place on an 18-limb projective point */
void ecc_double(fe p[18]);                          /* ecc.rs:186-246 */
void ecc_add(fe p[18], const fe q[18]);             /* ecc.rs:248-327 */
void ecc_add_mixed(fe p[18], const fe q_affine[12]); /* ecc.rs:329-404 */
/* constraint helpers; `result`, `cur`, `next` are 19 wide (point + bit) except reduce_x (18) */
void ecc_enforce_doubling(fe *result, const fe *cur, const fe *next, fe flag);                      /* ecc.rs:73-100 */
void ecc_enforce_addition_mixed(fe *result, const fe *cur, const fe *next, const fe *pt, fe flag);  /* ecc.rs:102-144 */
void ecc_enforce_addition_reduce_x(fe *result, const fe *cur, const fe *next, const fe *pt, fe flag); /* ecc.rs:146-172 */
#endif
