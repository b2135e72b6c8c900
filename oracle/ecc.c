/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Restates /root/reference/src/utils/ecc.rs (see ecc.h). */
#include "ecc.h"
#include "ref_constants.h"
#include <string.h>

static fe GEN[12], B3[6];
static int ready = 0;
void ecc_init_tables(void) {
    if (ready) return;
    for (int i = 0; i < 12; i++) GEN[i] = fe_from_u64(REF_GENERATOR[i]);
    for (int i = 0; i < 6; i++) B3[i] = fe_from_u64(REF_B3[i]);
    ready = 1;
}
const fe *ecc_generator(void) { ecc_init_tables(); return GEN; }

/* ---- Fp2: (a0 + a1 u), u^2 = 2u + 2 ---- */
typedef struct { fe c0, c1; } fp2;
static inline fp2 f2(const fe *a) { fp2 r = {a[0], a[1]}; return r; }
static inline fp2 f2_add(fp2 a, fp2 b) { fp2 r = {fe_add(a.c0, b.c0), fe_add(a.c1, b.c1)}; return r; }
static inline fp2 f2_sub(fp2 a, fp2 b) { fp2 r = {fe_sub(a.c0, b.c0), fe_sub(a.c1, b.c1)}; return r; }
static inline fp2 f2_dbl(fp2 a) { return f2_add(a, a); }
static inline fp2 f2_neg(fp2 a) { fp2 r = {fe_neg(a.c0), fe_neg(a.c1)}; return r; }
/* (a0 b0 + 2 a1 b1) + (a0 b1 + a1 b0 + 2 a1 b1) u   -- ecc.rs:424-439 */
static inline fp2 f2_mul(fp2 a, fp2 b) {
    fe p00 = fe_mul(a.c0, b.c0), p11 = fe_mul(a.c1, b.c1);
    fe cross = fe_mul(fe_sub(a.c0, a.c1), fe_sub(b.c1, b.c0)); /* a0b1 + a1b0 - p00 - p11 */
    fe c0 = fe_add(fe_dbl(p11), p00);
    fp2 r = {c0, fe_add(fe_add(p11, c0), cross)};
    return r;
}
static inline fp2 f2_sqr(fp2 a) { return f2_mul(a, a); } /* ecc.rs:407-421 is the same polynomial */
/* ecc.rs:441-446 */
static inline fp2 f2_inv(fp2 a) {
    fe t = fe_inv(fe_sub(fe_add(fe_sqr(a.c0), fe_mul(fe_dbl(a.c0), a.c1)), fe_dbl(fe_sqr(a.c1))));
    fp2 r = {fe_mul(fe_add(a.c0, fe_dbl(a.c1)), t), fe_mul(fe_neg(a.c1), t)};
    return r;
}

/* ---- Fp6: a + b v + c v^2 over Fp2, v^3 = -v - 1 ---- */
static void f6_store(fe r[6], fp2 a, fp2 b, fp2 c) { r[0] = a.c0; r[1] = a.c1; r[2] = b.c0; r[3] = b.c1; r[4] = c.c0; r[5] = c.c1; }
void fp6_mul(fe r[6], const fe x[6], const fe y[6]) {
    fp2 a = f2(x), b = f2(x + 2), c = f2(x + 4), d = f2(y), e = f2(y + 2), f = f2(y + 4);
    fp2 ad = f2_mul(a, d), be = f2_mul(b, e), cf = f2_mul(c, f);
    fp2 s_ab = f2_mul(f2_add(a, b), f2_add(d, e)); /* ad + be + (ae+bd) */
    fp2 s_ac = f2_mul(f2_add(a, c), f2_add(d, f));
    fp2 s_bc = f2_mul(f2_add(b, c), f2_add(e, f));
    fp2 sum = f2_add(f2_add(ad, be), cf);
    fp2 c0 = f2_sub(sum, s_bc);                                    /* ad - (bf+ce) */
    fp2 c1 = f2_sub(f2_sub(s_ab, s_bc), ad);                       /* (ae+bd) - (bf+ce) - cf */
    fp2 c2 = f2_add(f2_sub(f2_sub(s_ac, sum), cf), f2_dbl(be));    /* (af+cd) + be - cf */
    f6_store(r, c0, c1, c2);
}
void fp6_sqr(fe r[6], const fe a[6]) { fp6_mul(r, a, a); } /* ecc.rs:462-503 computes the same polynomial */
/* ecc.rs:551-591 */
void fp6_inv(fe r[6], const fe x[6]) {
    fp2 a = f2(x), b = f2(x + 2), c = f2(x + 4);
    fp2 a2 = f2_sqr(a), b2 = f2_sqr(b), c2 = f2_sqr(c);
    fp2 t = f2_mul(a, f2_add(a2, b2));
    t = f2_sub(t, f2_mul(b, b2));
    t = f2_add(t, f2_mul(f2_add(a, f2_sub(c, b)), c2));
    fp2 w = f2_mul(f2_add(f2_dbl(a), a), b);
    w = f2_mul(f2_sub(f2_dbl(a2), w), c);
    t = f2_inv(f2_sub(t, w));
    fp2 r0 = f2_add(f2_add(a2, b2), c2);
    r0 = f2_mul(f2_sub(r0, f2_mul(f2_sub(f2_dbl(a), b), c)), t);
    fp2 r1 = f2_mul(f2_neg(f2_add(f2_mul(a, b), c2)), t);
    fp2 r2 = f2_mul(f2_add(f2_sub(b2, f2_mul(a, c)), c2), t);
    f6_store(r, r0, r1, r2);
}
static void f6_add(fe r[6], const fe a[6], const fe b[6]) { for (int i = 0; i < 6; i++) r[i] = fe_add(a[i], b[i]); }
static void f6_sub(fe r[6], const fe a[6], const fe b[6]) { for (int i = 0; i < 6; i++) r[i] = fe_sub(a[i], b[i]); }
static void f6_dbl(fe r[6], const fe a[6]) { for (int i = 0; i < 6; i++) r[i] = fe_dbl(a[i]); }

/* RCB15 Algorithm 3 (exception-free doubling), a = 1 */
void ecc_double(fe p[18]) {
    ecc_init_tables();
    const fe *X = p, *Y = p + 6, *Z = p + 12;
    fe t0[6], t1[6], t2[6], t3[6], x3[6], y3[6], z3[6];
    fp6_sqr(t0, X); fp6_sqr(t1, Y); fp6_sqr(t2, Z);
    fp6_mul(t3, X, Y); f6_dbl(t3, t3);
    fp6_mul(z3, X, Z); f6_dbl(z3, z3);
    fp6_mul(y3, B3, t2); f6_add(y3, z3, y3);
    f6_sub(x3, t1, y3); f6_add(y3, t1, y3);
    fp6_mul(y3, x3, y3); fp6_mul(x3, t3, x3);
    fp6_mul(z3, B3, z3);
    f6_sub(t3, t0, t2); f6_add(t3, t3, z3);
    f6_dbl(z3, t0); f6_add(t0, z3, t0); f6_add(t0, t0, t2);
    fp6_mul(t0, t0, t3); f6_add(y3, y3, t0);
    fp6_mul(t2, Y, Z); f6_dbl(t2, t2);
    fp6_mul(t0, t2, t3); f6_sub(x3, x3, t0);
    fp6_mul(z3, t2, t1); f6_dbl(z3, z3); f6_dbl(z3, z3);
    memcpy(p, x3, 48); memcpy(p + 6, y3, 48); memcpy(p + 12, z3, 48);
}
/* RCB15 Algorithm 1 (complete addition), a = 1 */
void ecc_add(fe p[18], const fe q[18]) {
    ecc_init_tables();
    const fe *X1 = p, *Y1 = p + 6, *Z1 = p + 12, *X2 = q, *Y2 = q + 6, *Z2 = q + 12;
    fe t0[6], t1[6], t2[6], t3[6], t4[6], t5[6], x3[6], y3[6], z3[6];
    fp6_mul(t0, X1, X2); fp6_mul(t1, Y1, Y2); fp6_mul(t2, Z1, Z2);
    f6_add(t3, X1, Y1); f6_add(t4, X2, Y2); fp6_mul(t3, t3, t4);
    f6_add(t4, t0, t1); f6_sub(t3, t3, t4);
    f6_add(t4, X1, Z1); f6_add(t5, X2, Z2); fp6_mul(t4, t4, t5);
    f6_add(t5, t0, t2); f6_sub(t4, t4, t5);
    f6_add(t5, Y1, Z1); f6_add(x3, Y2, Z2); fp6_mul(t5, t5, x3);
    f6_add(x3, t1, t2); f6_sub(t5, t5, x3);
    fp6_mul(x3, B3, t2); f6_add(z3, x3, t4);
    f6_sub(x3, t1, z3); f6_add(z3, t1, z3);
    fp6_mul(y3, x3, z3);
    f6_dbl(t1, t0); f6_add(t1, t1, t0);
    fp6_mul(t4, B3, t4);
    f6_add(t1, t1, t2); f6_sub(t2, t0, t2);
    f6_add(t4, t4, t2);
    fp6_mul(t0, t1, t4); f6_add(y3, y3, t0);
    fp6_mul(t0, t5, t4); fp6_mul(x3, t3, x3); f6_sub(x3, x3, t0);
    fp6_mul(t0, t3, t1); fp6_mul(z3, t5, z3); f6_add(z3, z3, t0);
    memcpy(p, x3, 48); memcpy(p + 6, y3, 48); memcpy(p + 12, z3, 48);
}
/* RCB15 Algorithm 2 (complete mixed addition), a = 1 */
void ecc_add_mixed(fe p[18], const fe q[12]) {
    ecc_init_tables();
    const fe *X1 = p, *Y1 = p + 6, *Z1 = p + 12, *X2 = q, *Y2 = q + 6;
    fe t0[6], t1[6], t2[6], t3[6], t4[6], t5[6], x3[6], y3[6], z3[6];
    fp6_mul(t0, X1, X2); fp6_mul(t1, Y1, Y2);
    f6_add(t3, X2, Y2); f6_add(t4, X1, Y1); fp6_mul(t3, t3, t4);
    f6_add(t4, t0, t1); f6_sub(t3, t3, t4);
    fp6_mul(t4, X2, Z1); f6_add(t4, t4, X1);
    fp6_mul(t5, Y2, Z1); f6_add(t5, t5, Y1);
    fp6_mul(x3, Z1, B3); f6_add(z3, x3, t4);
    f6_sub(x3, t1, z3); f6_add(z3, t1, z3);
    fp6_mul(y3, x3, z3);
    f6_dbl(t1, t0); f6_add(t1, t1, t0);
    fp6_mul(t4, t4, B3);
    f6_add(t1, t1, Z1); f6_sub(t2, t0, Z1);
    f6_add(t4, t4, t2);
    fp6_mul(t0, t1, t4); f6_add(y3, y3, t0);
    fp6_mul(t0, t5, t4); fp6_mul(x3, t3, x3); f6_sub(x3, x3, t0);
    fp6_mul(t0, t3, t1); fp6_mul(z3, t5, z3); f6_add(z3, z3, t0);
    memcpy(p, x3, 48); memcpy(p + 6, y3, 48); memcpy(p + 12, z3, 48);
}

static inline void agg(fe *result, int i, fe flag, fe v) { result[i] = fe_add(result[i], fe_mul(flag, v)); }

void ecc_enforce_doubling(fe *result, const fe *cur, const fe *next, fe flag) {
    fe s1[18];
    memcpy(s1, cur, sizeof s1);
    ecc_double(s1);
    for (int i = 0; i < 18; i++) agg(result, i, flag, fe_sub(next[i], s1[i]));
    agg(result, 18, flag, fe_sub(fe_sqr(cur[18]), cur[18]));
}
void ecc_enforce_addition_mixed(fe *result, const fe *cur, const fe *next, const fe *pt, fe flag) {
    fe s1[18];
    memcpy(s1, cur, sizeof s1);
    ecc_add_mixed(s1, pt);
    fe bit = cur[18], nbit = fe_sub(FE_ONE, bit);
    for (int i = 0; i < 18; i++)
        agg(result, i, flag, fe_sub(next[i], fe_add(fe_mul(bit, s1[i]), fe_mul(nbit, cur[i]))));
    agg(result, 18, flag, fe_sub(cur[18], next[18]));
}
void ecc_enforce_addition_reduce_x(fe *result, const fe *cur, const fe *next, const fe *pt, fe flag) {
    fe s1[18], xz[6];
    memcpy(s1, cur, sizeof s1);
    ecc_add(s1, pt);
    fp6_mul(xz, next, s1 + 12);
    for (int i = 0; i < 6; i++) agg(result, i, flag, fe_sub(xz[i], s1[i]));
    for (int i = 6; i < 18; i++) agg(result, i, flag, fe_sub(next[i], s1[i]));
}
