// Curve arithmetic of the Schnorr sub-AIR (reference: src/utils/ecc.rs), shared by the host witness builder and
// the constraint kernels.  Fp2 = Fp[u]/(u^2 - 2u - 2), Fp6 = Fp2[v]/(v^3 + v + 1); curve y^2 = x^3 + x + B over
// Fp6 with the complete projective formulas of Renes-Costello-Batina (Alg. 1/2/3, a = 1, b3 = 3B).
// The AIR compares coordinates, not projective classes, so these exact polynomial formulas are part of the spec.
#pragma once
#include "field.cuh"
#include "ref_constants.h"

namespace ecc {
using f63::fe;

struct fp2 { fe c0, c1; };
struct fp6 { fe c[6]; };
struct point { fp6 x, y, z; };

CSG_HD fp2 add(fp2 a, fp2 b) { return {f63::add(a.c0, b.c0), f63::add(a.c1, b.c1)}; }
CSG_HD fp2 sub(fp2 a, fp2 b) { return {f63::sub(a.c0, b.c0), f63::sub(a.c1, b.c1)}; }
CSG_HD fp2 dbl(fp2 a) { return {f63::dbl(a.c0), f63::dbl(a.c1)}; }
CSG_HD fp2 neg(fp2 a) { return {f63::neg(a.c0), f63::neg(a.c1)}; }
// (a0 b0 + 2 a1 b1) + (a0 b1 + a1 b0 + 2 a1 b1) u : three base multiplications (ecc.rs:424-439)
CSG_HD fp2 mul(fp2 a, fp2 b) {
    fe p00 = f63::mul(a.c0, b.c0), p11 = f63::mul(a.c1, b.c1);
    fe cross = f63::mul(f63::sub(a.c0, a.c1), f63::sub(b.c1, b.c0));
    fe c0 = f63::add(f63::dbl(p11), p00);
    return {c0, f63::add(f63::add(p11, c0), cross)};
}
CSG_HD fp2 sqr(fp2 a) { return mul(a, a); }

CSG_HD fp6 add(const fp6 &a, const fp6 &b) { fp6 r; for (int i = 0; i < 6; i++) r.c[i] = f63::add(a.c[i], b.c[i]); return r; }
CSG_HD fp6 sub(const fp6 &a, const fp6 &b) { fp6 r; for (int i = 0; i < 6; i++) r.c[i] = f63::sub(a.c[i], b.c[i]); return r; }
CSG_HD fp6 dbl(const fp6 &a) { fp6 r; for (int i = 0; i < 6; i++) r.c[i] = f63::dbl(a.c[i]); return r; }
// (a + b v + c v^2)(d + e v + f v^2) with v^3 = -v - 1 (ecc.rs:506-548):
//     r0 = ad - bf - ce        r1 = ae + bd - bf - ce - cf        r2 = af + be + cd - cf
// Schoolbook over Fp with lazily reduced 128-bit sums: an Fp2 product (x0 + x1 u)(y0 + y1 u) is
//     (x0 y0 + x1 (2 y1))  +  (x0 y1 + x1 (y0 + 2 y1)) u
// so with 2 y1 and y0 + 2 y1 prepared per right operand it is four multiply-accumulates, and every output coefficient is ONE
// Montgomery reduction of a sum of at most 10 products (15 fit in 128 bits).  The subtractions become additions of products
// with p - b, p - c.  36 multiply-accumulates + 6 reductions instead of 18 modular multiplications + ~70 modular
// additions: half the ALU-pipe instructions, and the work is split evenly between the ALU and FMA pipes (IMAD.WIDE).
// Not inlined on the device: the curve formulas call it ~40 times per row.
struct fp2acc {
    f63::acc128 c0, c1;
    // += (x0 + x1 u) * y, y given as (y0, y1, 2 y1, y0 + 2 y1)
    CSG_HD void mac(fe x0, fe x1, fe y0, fe y1, fe y1d, fe ys) { c0.mac(x0, y0); c0.mac(x1, y1d); c1.mac(x0, y1); c1.mac(x1, ys); }
    CSG_HD void add(const fp2acc &o) { c0.add(o.c0); c1.add(o.c1); }
};
#if defined(__CUDACC__)
static __host__ __device__ __noinline__ fp6 mul(fp6 x, fp6 y) {   // by value: operands travel in registers, not through local memory
#else
inline fp6 mul(const fp6 &x, const fp6 &y) {
#endif
    const fe a0 = x.c[0], a1 = x.c[1], b0 = x.c[2], b1 = x.c[3], c0 = x.c[4], c1 = x.c[5];
    // p - b, p - c: operands only need to be below 2^64 and at most p for the product bound; p - 0 = p is fine
    const fe nb0 = f63::P - b0, nb1 = f63::P - b1, nc0 = f63::P - c0, nc1 = f63::P - c1;
    const fe d0 = y.c[0], d1 = y.c[1], e0 = y.c[2], e1 = y.c[3], f0 = y.c[4], f1 = y.c[5];
    const fe d1d = f63::dbl(d1), e1d = f63::dbl(e1), f1d = f63::dbl(f1);
    const fe ds = f63::add(d0, d1d), es = f63::add(e0, e1d), fs = f63::add(f0, f1d);
    fp2acc t;                                   // -(bf + ce), shared by r0 and r1
    t.mac(nb0, nb1, f0, f1, f1d, fs);
    t.mac(nc0, nc1, e0, e1, e1d, es);
    fp2acc r0 = t;
    r0.mac(a0, a1, d0, d1, d1d, ds);
    fp2acc r1 = t;
    r1.mac(a0, a1, e0, e1, e1d, es);
    r1.mac(b0, b1, d0, d1, d1d, ds);
    fp2acc u;                                   // -cf, shared by r1 and r2
    u.mac(nc0, nc1, f0, f1, f1d, fs);
    r1.add(u);
    fp2acc r2 = u;
    r2.mac(a0, a1, f0, f1, f1d, fs);
    r2.mac(b0, b1, e0, e1, e1d, es);
    r2.mac(c0, c1, d0, d1, d1d, ds);
    return {{r0.c0.reduce(), r0.c1.reduce(), r1.c0.reduce(), r1.c1.reduce(), r2.c0.reduce(), r2.c1.reduce()}};
}
// (a + b v + c v^2)^2 = (a^2 - 2bc) + (2ab - 2bc - c^2) v + (2ac + b^2 - c^2) v^2: six Fp2 products of which three are squares
// ((x0 + x1 u)^2 = x0^2 + 2 x1^2 + 2 x1 (x0 + x1) u: three multiply-accumulates instead of four) -- 21 multiply-accumulates and 6
// reductions against the 36 of mul(a, a).  The multiply-accumulate is what bounds the curve kernels (4 IMAD.WIDE = 16 cycles of
// the fmaheavy pipe each); a doubling has three squarings among its 13 products.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__ fp6 sqr(fp6 x) {
#else
inline fp6 sqr(const fp6 &x) {
#endif
    const fe a0 = x.c[0], a1 = x.c[1], b0 = x.c[2], b1 = x.c[3], c0 = x.c[4], c1 = x.c[5];
    const fe nb0 = f63::P - b0, nb1 = f63::P - b1, nc0 = f63::P - c0, nc1 = f63::P - c1;
    // right operands 2b and 2c of the cross products, as (y0, y1, 2 y1, y0 + 2 y1)
    const fe tb0 = f63::dbl(b0), tb1 = f63::dbl(b1), tb1d = f63::dbl(tb1), tbs = f63::add(tb0, tb1d);
    const fe tc0 = f63::dbl(c0), tc1 = f63::dbl(c1), tc1d = f63::dbl(tc1), tcs = f63::add(tc0, tc1d);
    fp2acc t;                                   // -2bc, shared by r0 and r1
    t.mac(nb0, nb1, tc0, tc1, tc1d, tcs);
    fp2acc u;                                   // -c^2, shared by r1 and r2
    u.c0.mac(nc0, c0); u.c0.mac(nc1, tc1); u.c1.mac(nc1, f63::dbl(f63::add(c0, c1)));
    fp2acc r0 = t;                              // + a^2
    r0.c0.mac(a0, a0); r0.c0.mac(a1, f63::dbl(a1)); r0.c1.mac(a1, f63::dbl(f63::add(a0, a1)));
    fp2acc r1 = t;                              // + 2ab - c^2
    r1.mac(a0, a1, tb0, tb1, tb1d, tbs);
    r1.add(u);
    fp2acc r2 = u;                              // + 2ac + b^2
    r2.mac(a0, a1, tc0, tc1, tc1d, tcs);
    r2.c0.mac(b0, b0); r2.c0.mac(b1, tb1); r2.c1.mac(b1, f63::dbl(f63::add(b0, b1)));
    return {{r0.c0.reduce(), r0.c1.reduce(), r1.c0.reduce(), r1.c1.reduce(), r2.c0.reduce(), r2.c1.reduce()}};
}
CSG_HD fp6 b3() { const uint64_t *t = CSG_TABLE(CSG_B3); return {{t[0], t[1], t[2], t[3], t[4], t[5]}}; }
CSG_HD fp6 load6(const fe *p) { return {{p[0], p[1], p[2], p[3], p[4], p[5]}}; }

// RCB15 Alg. 3 (ecc.rs:186-246)
CSG_HD point double_point(const point &p) {
    const fp6 B3 = b3();
    fp6 t0 = sqr(p.x), t1 = sqr(p.y), t2 = sqr(p.z);
    fp6 t3 = dbl(mul(p.x, p.y));
    fp6 z3 = dbl(mul(p.x, p.z));
    fp6 y3 = add(z3, mul(B3, t2));
    fp6 x3 = sub(t1, y3);
    y3 = add(t1, y3);
    y3 = mul(x3, y3);
    x3 = mul(t3, x3);
    z3 = mul(B3, z3);
    t3 = add(sub(t0, t2), z3);
    t0 = add(add(dbl(t0), t0), t2);
    t0 = mul(t0, t3);
    y3 = add(y3, t0);
    t2 = dbl(mul(p.y, p.z));
    x3 = sub(x3, mul(t2, t3));
    z3 = dbl(dbl(mul(t2, t1)));
    return {x3, y3, z3};
}
// RCB15 Alg. 2: projective + affine (ecc.rs:329-404)
CSG_HD point add_mixed(const point &p, const fp6 &qx, const fp6 &qy) {
    const fp6 B3 = b3();
    fp6 t0 = mul(p.x, qx), t1 = mul(p.y, qy);
    fp6 t3 = sub(mul(add(qx, qy), add(p.x, p.y)), add(t0, t1));
    fp6 t4 = add(mul(qx, p.z), p.x);
    fp6 t5 = add(mul(qy, p.z), p.y);
    fp6 z3 = add(mul(p.z, B3), t4);
    fp6 x3 = sub(t1, z3);
    z3 = add(t1, z3);
    fp6 y3 = mul(x3, z3);
    t1 = add(add(dbl(t0), t0), p.z);
    t4 = add(mul(t4, B3), sub(t0, p.z));
    y3 = add(y3, mul(t1, t4));
    x3 = sub(mul(t3, x3), mul(t5, t4));
    z3 = add(mul(t5, z3), mul(t3, t1));
    return {x3, y3, z3};
}
// RCB15 Alg. 1: projective + projective (ecc.rs:248-327)
CSG_HD point add_full(const point &p, const point &q) {
    const fp6 B3 = b3();
    fp6 t0 = mul(p.x, q.x), t1 = mul(p.y, q.y), t2 = mul(p.z, q.z);
    fp6 t3 = sub(mul(add(p.x, p.y), add(q.x, q.y)), add(t0, t1));
    fp6 t4 = sub(mul(add(p.x, p.z), add(q.x, q.z)), add(t0, t2));
    fp6 t5 = sub(mul(add(p.y, p.z), add(q.y, q.z)), add(t1, t2));
    fp6 z3 = add(mul(B3, t2), t4);
    fp6 x3 = sub(t1, z3);
    z3 = add(t1, z3);
    fp6 y3 = mul(x3, z3);
    t1 = add(add(dbl(t0), t0), t2);
    t4 = add(mul(B3, t4), sub(t0, t2));
    y3 = add(y3, mul(t1, t4));
    x3 = sub(mul(t3, x3), mul(t5, t4));
    z3 = add(mul(t5, z3), mul(t3, t1));
    return {x3, y3, z3};
}

// inversions for the final X/Z reduction of the witness (ecc.rs:441-446, 551-591)
CSG_HD fp2 inv(fp2 a) {
    fe t = f63::inv(f63::sub(f63::add(f63::sqr(a.c0), f63::mul(f63::dbl(a.c0), a.c1)), f63::dbl(f63::sqr(a.c1))));
    return {f63::mul(f63::add(a.c0, f63::dbl(a.c1)), t), f63::mul(f63::neg(a.c1), t)};
}
CSG_HD fp6 inv(const fp6 &x) {
    fp2 a = {x.c[0], x.c[1]}, b = {x.c[2], x.c[3]}, c = {x.c[4], x.c[5]};
    fp2 a2 = sqr(a), b2 = sqr(b), c2 = sqr(c);
    fp2 t = sub(mul(a, add(a2, b2)), mul(b, b2));
    t = add(t, mul(add(a, sub(c, b)), c2));
    fp2 w = mul(sub(dbl(a2), mul(add(dbl(a), a), b)), c);
    t = inv(sub(t, w));
    fp2 r0 = mul(sub(add(add(a2, b2), c2), mul(sub(dbl(a), b), c)), t);
    fp2 r1 = mul(neg(add(mul(a, b), c2)), t);
    fp2 r2 = mul(add(sub(b2, mul(a, c)), c2), t);
    return {{r0.c0, r0.c1, r1.c0, r1.c1, r2.c0, r2.c1}};
}
}  // namespace ecc
