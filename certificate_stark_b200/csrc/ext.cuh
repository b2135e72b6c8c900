// Extension fields of f63 for FieldExtension::Quadratic / Cubic (ProofOptions, /root/reference/src/lib.rs:83; the reference's
// tests prove and verify with None, Quadratic and Cubic, src/tests.rs:12-30, and its example binary defaults to Cubic,
// examples/state-transition.rs:62-66).  Shared by the host driver, the host verifier and the kernels.
//
// The `ExtensibleField<2>/<3>` implementations for f63 live in the un-vendored winterfell fork; the polynomials here are the
// ones the reference tree uses for its own curve tower (src/utils/ecc.rs:424-548), both irreducible over this prime:
//     d = 2:  u^2 = 2u + 2          d = 3:  v^3 = -v - 1
// An element is d base elements (Montgomery form); on the device E-valued vectors are stored as d planes of base elements,
// so every transform of the base field (NTT, LDE, linear combinations) applies plane by plane.
#pragma once
#include "field.cuh"

namespace f63 {

struct xe { fe c[3]; };   // components beyond the degree stay zero

CSG_HD xe x_from(fe a) { xe r; r.c[0] = a; r.c[1] = 0; r.c[2] = 0; return r; }
CSG_HD xe x_zero() { return x_from(0); }
CSG_HD xe x_one() { return x_from(ONE); }
CSG_HD bool x_eq(const xe &a, const xe &b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2]; }
CSG_HD xe x_add(const xe &a, const xe &b) { xe r; for (int j = 0; j < 3; j++) r.c[j] = add(a.c[j], b.c[j]); return r; }
CSG_HD xe x_sub(const xe &a, const xe &b) { xe r; for (int j = 0; j < 3; j++) r.c[j] = sub(a.c[j], b.c[j]); return r; }
CSG_HD xe x_scale(const xe &a, fe s) { xe r; for (int j = 0; j < 3; j++) r.c[j] = mul(a.c[j], s); return r; }
CSG_HD xe x_add_base(xe a, fe b) { a.c[0] = add(a.c[0], b); return a; }
CSG_HD xe x_mul(int d, const xe &a, const xe &b) {
    xe r = x_zero();
    if (d == 1) { r.c[0] = mul(a.c[0], b.c[0]); return r; }
    if (d == 2) {
        const fe t = dbl(mul(a.c[1], b.c[1]));
        r.c[0] = add(mul(a.c[0], b.c[0]), t);
        r.c[1] = add(add(mul(a.c[0], b.c[1]), mul(a.c[1], b.c[0])), t);
        return r;
    }
    acc128 s1, s2, s3;
    s1.mac(a.c[0], b.c[1]); s1.mac(a.c[1], b.c[0]);
    s2.mac(a.c[0], b.c[2]); s2.mac(a.c[1], b.c[1]); s2.mac(a.c[2], b.c[0]);
    s3.mac(a.c[1], b.c[2]); s3.mac(a.c[2], b.c[1]);
    const fe p0 = mul(a.c[0], b.c[0]), p1 = s1.reduce(), p2 = s2.reduce(), p3 = s3.reduce(), p4 = mul(a.c[2], b.c[2]);
    r.c[0] = sub(p0, p3);             // v^3 = -v - 1
    r.c[1] = sub(sub(p1, p3), p4);    // v^4 = -v^2 - v
    r.c[2] = sub(p2, p4);
    return r;
}
CSG_HD xe x_pow(int d, xe b, uint64_t e) {
    xe r = x_one();
    while (e) { if (e & 1) r = x_mul(d, r, b); b = x_mul(d, b, b); e >>= 1; }
    return r;
}
// Frobenius a -> a^p = a0 + a1 * phi^p + a2 * phi^(2p); the two constants of the cubic field travel in ExtConsts
struct ExtConsts { xe f1, f2; };
CSG_HD xe x_frobenius(int d, const xe &a, const ExtConsts &k) {
    if (d == 1) return a;
    if (d == 2) { xe r = x_zero(); r.c[0] = add(a.c[0], dbl(a.c[1])); r.c[1] = neg(a.c[1]); return r; }   // conj(u) = 2 - u
    return x_add(x_from(a.c[0]), x_add(x_scale(k.f1, a.c[1]), x_scale(k.f2, a.c[2])));
}
// inverse through the norm: a^-1 = (product of the other conjugates) / N(a), N(a) in the base field; inv(0) = 0
CSG_HD xe x_inv(int d, const xe &a, const ExtConsts &k) {
    if (d == 1) return x_from(inv(a.c[0]));
    xe t = x_frobenius(d, a, k);
    if (d == 3) t = x_mul(3, t, x_frobenius(3, t, k));
    const xe n = x_mul(d, a, t);
    return x_scale(t, inv(n.c[0]));
}
inline ExtConsts ext_consts() {
    ExtConsts k;
    xe v = x_zero(); v.c[1] = ONE;
    k.f1 = x_pow(3, v, P);
    k.f2 = x_mul(3, k.f1, k.f1);
    return k;
}

}  // namespace f63
