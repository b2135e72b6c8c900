/* ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
 *
 * Extension fields of f63 for FieldExtension::Quadratic / Cubic (ProofOptions at src/lib.rs:83; the reference's tests
 * prove and verify with all three settings, src/tests.rs:12-30, and its example binary defaults to Cubic,
 * examples/state-transition.rs:62-66).  The `ExtensibleField<2>` / `<3>` implementations for f63 live in the un-vendored
 * winterfell fork (Cargo.toml:20); the polynomials used here are the ones the reference tree itself uses for its curve
 * tower (src/utils/ecc.rs:424-548):
 *     quadratic  E2 = Fp[u] / (u^2 - 2u - 2)      (12 is a non-residue mod p)
 *     cubic      E3 = Fp[v] / (v^3 + v + 1)        (no root mod p)
 * upstream winterfell's x^2 - x - 1 and x^3 - x + 2 (f62/f64) are reducible over this prime.
 * An element is d base elements c[0..d), serialised in that order as canonical little-endian words. */
#ifndef ORACLE_EXT_H
#define ORACLE_EXT_H
#include "f63.h"

typedef struct { fe c[3]; } xe; /* components beyond the degree are zero */

static inline xe xe_from_fe(fe a) { xe r = {{a, 0, 0}}; return r; }
static inline xe xe_zero(void) { xe r = {{0, 0, 0}}; return r; }
static inline xe xe_one(void) { return xe_from_fe(FE_ONE); }
static inline int xe_eq(xe a, xe b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2]; }
static inline xe xe_add(xe a, xe b) { xe r = {{fe_add(a.c[0], b.c[0]), fe_add(a.c[1], b.c[1]), fe_add(a.c[2], b.c[2])}}; return r; }
static inline xe xe_sub(xe a, xe b) { xe r = {{fe_sub(a.c[0], b.c[0]), fe_sub(a.c[1], b.c[1]), fe_sub(a.c[2], b.c[2])}}; return r; }
static inline xe xe_scale(xe a, fe s) { xe r = {{fe_mul(a.c[0], s), fe_mul(a.c[1], s), fe_mul(a.c[2], s)}}; return r; }
/* d = 1: base field; d = 2: u^2 = 2u + 2; d = 3: v^3 = -v - 1 */
static inline xe xe_mul(int d, xe a, xe b) {
    xe r = {{0, 0, 0}};
    if (d == 1) { r.c[0] = fe_mul(a.c[0], b.c[0]); return r; }
    if (d == 2) {
        fe t = fe_dbl(fe_mul(a.c[1], b.c[1]));
        r.c[0] = fe_add(fe_mul(a.c[0], b.c[0]), t);
        r.c[1] = fe_add(fe_add(fe_mul(a.c[0], b.c[1]), fe_mul(a.c[1], b.c[0])), t);
        return r;
    }
    fe p0 = fe_mul(a.c[0], b.c[0]);
    fe p1 = fe_add(fe_mul(a.c[0], b.c[1]), fe_mul(a.c[1], b.c[0]));
    fe p2 = fe_add(fe_add(fe_mul(a.c[0], b.c[2]), fe_mul(a.c[1], b.c[1])), fe_mul(a.c[2], b.c[0]));
    fe p3 = fe_add(fe_mul(a.c[1], b.c[2]), fe_mul(a.c[2], b.c[1]));
    fe p4 = fe_mul(a.c[2], b.c[2]);
    r.c[0] = fe_sub(p0, p3);                 /* v^3 = -v - 1 */
    r.c[1] = fe_sub(fe_sub(p1, p3), p4);     /* v^4 = -v^2 - v */
    r.c[2] = fe_sub(p2, p4);
    return r;
}
static inline xe xe_exp(int d, xe b, uint64_t e) {
    xe r = xe_one();
    while (e) { if (e & 1) r = xe_mul(d, r, b); b = xe_mul(d, b, b); e >>= 1; }
    return r;
}
/* Frobenius a -> a^p: a0 + a1 * phi^p + a2 * phi^(2p) */
static inline xe xe_frobenius(int d, xe a) {
    if (d == 1) return a;
    if (d == 2) { xe r = {{fe_add(a.c[0], fe_dbl(a.c[1])), fe_neg(a.c[1]), 0}}; return r; }   /* conj(u) = 2 - u */
    static xe f1, f2; static int ready = 0;
    if (!ready) {
        xe v = {{0, FE_ONE, 0}};
        xe t = xe_exp(3, v, F63_P);
        f1 = t; f2 = xe_mul(3, t, t);
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    return xe_add(xe_from_fe(a.c[0]), xe_add(xe_scale(f1, a.c[1]), xe_scale(f2, a.c[2])));
}
/* inverse through the norm: a^-1 = (prod of the other conjugates) / N(a); inv(0) = 0 */
static inline xe xe_inv(int d, xe a) {
    if (d == 1) return xe_from_fe(fe_inv(a.c[0]));
    xe t = xe_frobenius(d, a);
    if (d == 3) t = xe_mul(3, t, xe_frobenius(3, t));   /* a^p * a^(p^2) */
    xe n = xe_mul(d, a, t);                             /* the norm: lies in the base field */
    return xe_scale(t, fe_inv(n.c[0]));
}
#endif
