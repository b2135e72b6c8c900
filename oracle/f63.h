/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * f63: the prime field of winterfell::math::fields::f63::BaseElement used by the reference
 * (src/prover.rs:2, src/air.rs:41).  The type lives in the un-vendored ToposWare winterfell fork
 * (Cargo.toml:20, rev 8e37310); what the reference tree itself pins is
 *   - the modulus M = 4719772409484279809 = 2^62 + 2^56 + 2^55 + 1   (src/range/tests.rs:59, benches/range.rs:23)
 *   - Montgomery representation with R = 2^64: GENERATOR's from_raw_unchecked limbs lie on the curve only
 *     after multiplying by 2^-64 (src/utils/ecc.rs:23-45)
 *   - BaseElement::new(v) takes canonical v (MDS * INV_MDS = I holds that way, src/utils/rescue.rs:385-781)
 *   - to_bytes() = 8-byte canonical little-endian (src/range/prover.rs:29-30, bits are read Lsb0)
 * Elements are kept in Montgomery form, like the reference does.
 */
#ifndef ORACLE_F63_H
#define ORACLE_F63_H
#include <stdint.h>
#include <stddef.h>

typedef uint64_t fe; /* Montgomery form, always reduced to [0, p) */
typedef unsigned __int128 u128;

#define F63_P 0x4180000000000001ULL
#define F63_NPRIME 0x417fffffffffffffULL /* -p^-1 mod 2^64 */
#define F63_R 0x3b7ffffffffffffdULL      /* 2^64 mod p  == Montgomery(1) */
#define F63_R2 0x32734c36b7b1d512ULL     /* 2^128 mod p */
#define F63_TWO_ADICITY 55
/* multiplicative generator: smallest primitive root of p.  The fork's choice is not visible from the
 * reference tree; 3 is what upstream winterfell uses for f62/f128 and is the smallest primitive root here. */
#define F63_GENERATOR 3ULL
/* 3^131 mod p: a primitive 2^55-th root of unity (canonical) */
#define F63_TWO_ADIC_ROOT 0x0141727b75b35c50ULL

#define FE_ZERO ((fe)0)
#define FE_ONE ((fe)F63_R)

static inline fe fe_redc(u128 t) {
    uint64_t m = (uint64_t)t * F63_NPRIME;
    u128 mp = (u128)m * F63_P;
    /* (t + m*p) / 2^64 ; t < p*2^64 so the result is < 2p */
    uint64_t lo = (uint64_t)t, hi = (uint64_t)(t >> 64);
    uint64_t carry = (lo != 0); /* low words sum to exactly 2^64 unless lo == 0 */
    uint64_t u = hi + (uint64_t)(mp >> 64) + carry;
    return u >= F63_P ? u - F63_P : u;
}
static inline fe fe_mul(fe a, fe b) { return fe_redc((u128)a * b); }
static inline fe fe_sqr(fe a) { return fe_mul(a, a); }
static inline fe fe_add(fe a, fe b) { uint64_t s = a + b; return s >= F63_P ? s - F63_P : s; }
static inline fe fe_sub(fe a, fe b) { return a >= b ? a - b : a + F63_P - b; }
static inline fe fe_neg(fe a) { return a ? F63_P - a : 0; }
static inline fe fe_dbl(fe a) { return fe_add(a, a); }
/* canonical integer (any u64, reduced mod p) -> Montgomery */
static inline fe fe_from_u64(uint64_t v) { return fe_mul(v % F63_P, F63_R2); }
/* Montgomery -> canonical */
static inline uint64_t fe_to_u64(fe a) { return fe_redc((u128)a); }
static inline fe fe_exp(fe b, uint64_t e) {
    fe r = FE_ONE;
    while (e) { if (e & 1) r = fe_mul(r, b); b = fe_sqr(b); e >>= 1; }
    return r;
}
static inline fe fe_inv(fe a) { return fe_exp(a, F63_P - 2); } /* inv(0) = 0 like winterfell */
/* primitive 2^logn-th root of unity, Montgomery form (winterfell StarkField::get_root_of_unity) */
static inline fe fe_root_of_unity(unsigned logn) {
    fe r = fe_from_u64(F63_TWO_ADIC_ROOT);
    for (unsigned i = logn; i < F63_TWO_ADICITY; i++) r = fe_sqr(r);
    return r;
}
static inline unsigned ilog2(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) l++; return l; }
#endif
