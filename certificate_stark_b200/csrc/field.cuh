// f63 arithmetic shared by the host driver (C++) and the sm_100a kernels.
//
// The field is winterfell::math::fields::f63::BaseElement of the reference (src/prover.rs:2, src/air.rs:41):
//   p = 2^62 + 2^56 + 2^55 + 1 = 0x4180000000000001   (reference: src/range/tests.rs:59, benches/range.rs:23)
// Elements are kept in Montgomery form (R = 2^64) end to end on the device, as the reference does on the CPU
// (src/utils/ecc.rs:23-36 stores GENERATOR as raw Montgomery limbs); bytes that get hashed or serialised are
// converted to canonical little-endian first.
//
// p and -p^-1 = p - 2 are both sparse, so the Montgomery reduction needs no multiplier: on the device it is
// shifts and adds on the ALU pipe, leaving the fma pipe for the 4 IMAD.WIDE of the 64x64 product.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define CSG_HD __host__ __device__ __forceinline__
#define CSG_D __device__ __forceinline__
#else
#define CSG_HD inline
#endif

namespace f63 {

typedef uint64_t fe;  // Montgomery form, reduced to [0, p)

constexpr uint64_t P = 0x4180000000000001ULL;
constexpr uint64_t NPRIME = 0x417fffffffffffffULL;  // -p^-1 mod 2^64 == p - 2
constexpr uint64_t R = 0x3b7ffffffffffffdULL;       // 2^64 mod p  (Montgomery form of 1)
constexpr uint64_t R2 = 0x32734c36b7b1d512ULL;      // 2^128 mod p
constexpr unsigned TWO_ADICITY = 55;
constexpr uint64_t GENERATOR = 3;                          // multiplicative generator / LDE domain offset (canonical)
constexpr uint64_t TWO_ADIC_ROOT = 0x0141727b75b35c50ULL;  // 3^131 mod p, canonical
constexpr fe ZERO = 0, ONE = R;

struct u128 { uint64_t lo, hi; };
#if defined(CSG_REDC_CHECK) && defined(__CUDACC__)
static __device__ unsigned long long csg_redc_violations = 0;
#endif

CSG_HD u128 mul_wide(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return {a * b, __umul64hi(a, b)};
#else
    unsigned __int128 t = (unsigned __int128)a * b;
    return {(uint64_t)t, (uint64_t)(t >> 64)};
#endif
}

// Montgomery reduction of t = hi * 2^64 + lo, t < p * 2^64.  Returns t * 2^-64 mod p in [0, p).
// p = P_HI * 2^32 + 1, so -p^-1 mod 2^32 = 2^32 - 1 and a 32-bit reduction step is:  m = -t0 (mod 2^32);
// t <- (t + m*p) / 2^32 = (t >> 32) + (t0 != 0) + m * P_HI.  Two steps take the 128-bit product down to < 2p with
// two 32x32->64 multiply-adds (IMAD.WIDE.U32 on the device) and a handful of adds: no 64-bit multiplications at all.
constexpr uint64_t P_HI = P >> 32;   // 0x41800000
CSG_HD fe redc_reference(uint64_t lo, uint64_t hi);
// the reduction without its final conditional subtraction: t * 2^-64 mod p as a value in [0, t / 2^64 + p], i.e. below 2p
// whenever t < p * 2^64 (lazy butterflies of the NTT keep their operands in [0, 2p))
CSG_HD uint64_t redc_raw(uint64_t lo, uint64_t hi) {
#if defined(CSG_REDC_REFERENCE)
    return redc_reference(lo, hi);
#endif
#if defined(CSG_REDC_CHECK) && defined(__CUDA_ARCH__)
    if (hi >= P) atomicAdd(&csg_redc_violations, 1ULL);
#endif
#if defined(__CUDA_ARCH__)
    // One asm block: the carry flag links the steps (t0 + m1 carries exactly when t0 != 0), and the sequence stays
    // opaque to the optimiser -- the plain C++ below was observed to be mis-optimised inside one large kernel
    // (tests/test_gpu_parity.py::test_schnorr_proof_identical_to_oracle caught it).
    uint64_t u;
    asm("{\n\t"
        ".reg .u32 t0, l1, m, pl, ph, t1, r0, r1, h0, h1, d;\n\t"
        ".reg .u64 mp;\n\t"
        "mov.b64 {t0, l1}, %1;\n\t"
        "mov.b64 {h0, h1}, %2;\n\t"
        "sub.u32 m, 0, t0;\n\t"
        "mul.wide.u32 mp, m, 0x41800000;\n\t"
        "mov.b64 {pl, ph}, mp;\n\t"
        "add.cc.u32 d, t0, m;\n\t"          // carry = (t0 != 0)
        "addc.cc.u32 t1, l1, pl;\n\t"       // word 1 of t + m*p
        "addc.cc.u32 r0, h0, ph;\n\t"       // rest = hi + (m*P_HI >> 32) + carry
        "addc.u32 r1, h1, 0;\n\t"
        "sub.u32 m, 0, t1;\n\t"
        "mul.wide.u32 mp, m, 0x41800000;\n\t"
        "mov.b64 {pl, ph}, mp;\n\t"
        "add.cc.u32 d, t1, m;\n\t"          // carry = (t1 != 0)
        "addc.cc.u32 r0, r0, pl;\n\t"
        "addc.u32 r1, r1, ph;\n\t"
        "mov.b64 %0, {r0, r1};\n\t"
        "}"
        : "=l"(u) : "l"(lo), "l"(hi));
    return u;
#else
    const uint32_t t0 = (uint32_t)lo, m1 = 0u - t0;
    const uint64_t mp1 = (uint64_t)m1 * P_HI;                                   // < 2^63
    const uint64_t s = (lo >> 32) + (uint64_t)(uint32_t)mp1 + (t0 != 0 ? 1 : 0);   // word 1 of t + m1*p, with its carry
    const uint32_t t1 = (uint32_t)s, m2 = 0u - t1;
    const uint64_t rest = hi + (mp1 >> 32) + (s >> 32);                         // (t + m1*p) >> 64, at most p
    const uint64_t u = (uint64_t)m2 * P_HI + rest + (t1 != 0 ? 1 : 0);           // < 2p
    return u;
#endif
}
CSG_HD fe redc(uint64_t lo, uint64_t hi) { const uint64_t u = redc_raw(lo, hi); return u >= P ? u - P : u; }
// the same value computed the long way (one 64-bit Montgomery step with -p^-1 = p - 2); kept for the unit tests
CSG_HD fe redc_reference(uint64_t lo, uint64_t hi) {
    uint64_t m = lo * NPRIME;
    u128 mp = mul_wide(m, P);
    uint64_t u = hi + mp.hi + (lo != 0 ? 1 : 0);
    return u >= P ? u - P : u;
}

// a * b * 2^-64 mod p as a value in [0, a * b / 2^64 + p] (below 2p when a * b < p * 2^64): product and reduction in one
// carry chain on the device.  The four partial products land directly in the words they belong to (two of them as
// multiply-adds onto the middle words), and each reduction step m = -t_k, t += m * p is ONE multiply-add whose carry-in is
// the carry of t_k + m (set exactly when t_k != 0): 22 SASS instructions for a full modular multiplication instead of 28.
CSG_HD uint64_t mul_raw(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__) && !defined(CSG_REDC_REFERENCE) && !defined(CSG_REDC_CHECK)
    uint64_t u;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, t0, t1, t2, t3, m, d;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "mul.lo.u32 t0, a0, b0;\n\t"
        "mul.hi.u32 t1, a0, b0;\n\t"
        "mul.lo.u32 t2, a1, b1;\n\t"
        "mul.hi.u32 t3, a1, b1;\n\t"
        "mad.lo.cc.u32 t1, a0, b1, t1;\n\t"
        "madc.hi.cc.u32 t2, a0, b1, t2;\n\t"
        "addc.u32 t3, t3, 0;\n\t"
        "mad.lo.cc.u32 t1, a1, b0, t1;\n\t"
        "madc.hi.cc.u32 t2, a1, b0, t2;\n\t"
        "addc.u32 t3, t3, 0;\n\t"
        "sub.u32 m, 0, t0;\n\t"
        "add.cc.u32 d, t0, m;\n\t"                     // carry = (t0 != 0)
        "madc.lo.cc.u32 t1, m, 0x41800000, t1;\n\t"
        "madc.hi.cc.u32 t2, m, 0x41800000, t2;\n\t"
        "addc.u32 t3, t3, 0;\n\t"
        "sub.u32 m, 0, t1;\n\t"
        "add.cc.u32 d, t1, m;\n\t"                     // carry = (t1 != 0)
        "madc.lo.cc.u32 t2, m, 0x41800000, t2;\n\t"
        "madc.hi.u32 t3, m, 0x41800000, t3;\n\t"
        "mov.b64 %0, {t2, t3};\n\t"
        "}"
        : "=l"(u) : "l"(a), "l"(b));
    return u;
#else
    u128 t = mul_wide(a, b);
    return redc_raw(t.lo, t.hi);
#endif
}
CSG_HD fe mul(fe a, fe b) { const uint64_t u = mul_raw(a, b); return u >= P ? u - P : u; }
CSG_HD fe sqr(fe a) { return mul(a, a); }
// lazy arithmetic on values in [0, 2p) (2p < 2^64; 4p is not, so sums are brought back below 2p at once):
//   mul_2p: a < 2p, b < p  ->  a*b*2^-64 mod p, below 1.52 p, no conditional subtraction
//   add_2p / sub_2p: operands and result in [0, 2p)
CSG_HD uint64_t mul_2p(uint64_t a, fe b) { return mul_raw(a, b); }
CSG_HD uint64_t add_2p(uint64_t a, uint64_t b) { uint64_t s = a + b; return s >= 2 * P ? s - 2 * P : s; }   // needs a + b < 2^64: one operand from mul_2p
CSG_HD uint64_t add_2p_any(uint64_t a, uint64_t b) { const uint64_t t = 2 * P - b; return a >= t ? a - t : a + b; }   // any a, b < 2p (4p > 2^64)
CSG_HD uint64_t sub_2p(uint64_t a, uint64_t b) { return a >= b ? a - b : a + 2 * P - b; }
CSG_HD fe reduce_2p(uint64_t a) { return a >= P ? a - P : a; }
CSG_HD fe add(fe a, fe b) { uint64_t s = a + b; return s >= P ? s - P : s; }
CSG_HD fe sub(fe a, fe b) { return a >= b ? a - b : a + P - b; }
CSG_HD fe neg(fe a) { return a ? P - a : 0; }
CSG_HD fe dbl(fe a) { return add(a, a); }
CSG_HD fe to_mont(uint64_t canonical) { return mul(canonical, R2); }   // canonical must be < p
CSG_HD uint64_t from_mont(fe a) { return redc(a, 0); }
CSG_HD fe pow(fe b, uint64_t e) {
    fe r = ONE;
    while (e) { if (e & 1) r = mul(r, b); b = sqr(b); e >>= 1; }
    return r;
}
CSG_HD fe inv(fe a) { return pow(a, P - 2); }

// ---- lazily reduced accumulators for sums of products.
// Device form: the four 32x32 partial products of a term go to three 64-bit COLUMNS that are never aligned with each other
// until the end -- e (weight 1) += a0 b0, o (weight 2^32) += a0 b1 + a1 b0, h (weight 2^64) += a1 b1 -- so a term is four
// IMAD.WIDE with the 64-bit addend built in, the carry out of column e rides into column h as the carry-in of its
// multiply-add, and the carries out of column o are counted (ptxas folds two counts into one IADD3.X).  5-6 instructions
// per term instead of the ~15 of a 128-bit multiply-add with full carry propagation, most of them on the FMA pipe.
// Host form: plain 64-bit words.

// sum of up to 14 products a_i * b_i (each < p^2) fits in 128 bits
struct acc128 {
#if defined(__CUDA_ARCH__)
    uint32_t e0, e1, o0, o1, h0, h1;   // value = e + o * 2^32 + h * 2^64 (the total is below 2^128, so h cannot overflow)
    CSG_HD acc128() : e0(0), e1(0), o0(0), o1(0), h0(0), h1(0) {}
    CSG_HD void mac(fe a, fe b) {
        asm("{\n\t"
            ".reg .u32 a0, a1, b0, b1;\n\t"
            "mov.b64 {a0, a1}, %6;\n\t"
            "mov.b64 {b0, b1}, %7;\n\t"
            "mad.lo.cc.u32 %0, a0, b0, %0;\n\t"
            "madc.hi.cc.u32 %1, a0, b0, %1;\n\t"
            "madc.lo.cc.u32 %4, a1, b1, %4;\n\t"    // carry of column e has weight 2^64: the low word of column h
            "madc.hi.u32 %5, a1, b1, %5;\n\t"
            "mad.lo.cc.u32 %2, a0, b1, %2;\n\t"
            "madc.hi.cc.u32 %3, a0, b1, %3;\n\t"
            "addc.u32 %5, %5, 0;\n\t"               // carry of column o has weight 2^96: the high word of column h
            "mad.lo.cc.u32 %2, a1, b0, %2;\n\t"
            "madc.hi.cc.u32 %3, a1, b0, %3;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "}"
            : "+r"(e0), "+r"(e1), "+r"(o0), "+r"(o1), "+r"(h0), "+r"(h1) : "l"(a), "l"(b));
    }
    CSG_HD void add(const acc128 &x) {
        asm("add.cc.u32 %0, %0, %6;\n\t"
            "addc.cc.u32 %1, %1, %7;\n\t"
            "addc.cc.u32 %4, %4, %10;\n\t"
            "addc.u32 %5, %5, %11;\n\t"
            "add.cc.u32 %2, %2, %8;\n\t"
            "addc.cc.u32 %3, %3, %9;\n\t"
            "addc.u32 %5, %5, 0;"
            : "+r"(e0), "+r"(e1), "+r"(o0), "+r"(o1), "+r"(h0), "+r"(h1) : "r"(x.e0), "r"(x.e1), "r"(x.o0), "r"(x.o1), "r"(x.h0), "r"(x.h1));
    }
    CSG_HD void words(uint64_t &lo, uint64_t &hi) const {
        asm("{\n\t"
            ".reg .u32 l1, l2, l3;\n\t"
            "add.cc.u32 l1, %3, %4;\n\t"
            "addc.cc.u32 l2, %6, %5;\n\t"
            "addc.u32 l3, %7, 0;\n\t"
            "mov.b64 %0, {%2, l1};\n\t"
            "mov.b64 %1, {l2, l3};\n\t"
            "}"
            : "=l"(lo), "=l"(hi) : "r"(e0), "r"(e1), "r"(o0), "r"(o1), "r"(h0), "r"(h1));
    }
#else
    uint64_t lo, hi;
    CSG_HD acc128() : lo(0), hi(0) {}
    CSG_HD void mac(fe a, fe b) {
        u128 t = mul_wide(a, b);
        lo += t.lo;
        hi += t.hi + (lo < t.lo ? 1 : 0);
    }
    CSG_HD void add(const acc128 &x) { lo += x.lo; hi += x.hi + (lo < x.lo ? 1 : 0); }
    CSG_HD void words(uint64_t &l, uint64_t &h) const { l = lo; h = hi; }
#endif
    // value mod p (Montgomery-reduced): bring hi below p first so that t < p * 2^64
    CSG_HD fe reduce() const {
        uint64_t l, h;
        words(l, h);
        if (h >= 2 * P) h -= 2 * P;
        if (h >= P) h -= P;
        return redc(l, h);
    }
};

// 192-bit accumulator as three plain words: the form in which partial sums are parked in memory (airs.cuh, split mode)
struct acc192w {
    uint64_t lo, mid, hi;
    CSG_HD acc192w() : lo(0), mid(0), hi(0) {}
    CSG_HD void mac(fe a, fe b) {
#if defined(__CUDA_ARCH__)
        asm("mad.lo.cc.u64 %0, %3, %4, %0;\n\tmadc.hi.cc.u64 %1, %3, %4, %1;\n\taddc.u64 %2, %2, 0;" : "+l"(lo), "+l"(mid), "+l"(hi) : "l"(a), "l"(b));
        return;
#endif
        u128 t = mul_wide(a, b);
        lo += t.lo;
        uint64_t c = lo < t.lo ? 1 : 0;
        uint64_t m = mid + t.hi;       // t.hi < 2^61, so adding the carry below cannot wrap a second time
        uint64_t c2 = m < mid ? 1 : 0;
        m += c;
        c2 += m < c ? 1 : 0;
        mid = m;
        hi += c2;
    }
    // (hi*2^128 + mid*2^64 + lo) * 2^-64 mod p  =  hi*2^64 + mid + lo*2^-64   (hi counts carries: far below p)
    CSG_HD fe reduce() const {
        uint64_t m = mid;
        if (m >= 2 * P) m -= 2 * P;
        if (m >= P) m -= P;
        return add(add(m, redc(lo, 0)), mul(hi, R2));
    }
};

// ---- lazily reduced 192-bit accumulator for long sums of products (the random linear combination of constraints)
#if defined(__CUDA_ARCH__)
struct acc192 {
    uint32_t e0, e1, o0, o1, h0, h1, c1, c2;   // value = e + o * 2^32 + h * 2^64 + c1 * 2^96 + c2 * 2^128
    CSG_HD acc192() : e0(0), e1(0), o0(0), o1(0), h0(0), h1(0), c1(0), c2(0) {}
    CSG_HD void mac(fe a, fe b) {
        asm("{\n\t"
            ".reg .u32 a0, a1, b0, b1;\n\t"
            "mov.b64 {a0, a1}, %8;\n\t"
            "mov.b64 {b0, b1}, %9;\n\t"
            "mad.lo.cc.u32 %0, a0, b0, %0;\n\t"
            "madc.hi.cc.u32 %1, a0, b0, %1;\n\t"
            "madc.lo.cc.u32 %4, a1, b1, %4;\n\t"
            "madc.hi.cc.u32 %5, a1, b1, %5;\n\t"
            "addc.u32 %7, %7, 0;\n\t"
            "mad.lo.cc.u32 %2, a0, b1, %2;\n\t"
            "madc.hi.cc.u32 %3, a0, b1, %3;\n\t"
            "addc.u32 %6, %6, 0;\n\t"
            "mad.lo.cc.u32 %2, a1, b0, %2;\n\t"
            "madc.hi.cc.u32 %3, a1, b0, %3;\n\t"
            "addc.u32 %6, %6, 0;\n\t"
            "}"
            : "+r"(e0), "+r"(e1), "+r"(o0), "+r"(o1), "+r"(h0), "+r"(h1), "+r"(c1), "+r"(c2) : "l"(a), "l"(b));
    }
    CSG_HD fe reduce() const {
        acc192w w;
        uint32_t top;
        asm("{\n\t"
            ".reg .u32 l1, m0, m1;\n\t"
            "add.cc.u32 l1, %4, %5;\n\t"
            "addc.cc.u32 m0, %7, %6;\n\t"
            "addc.cc.u32 m1, %8, %9;\n\t"
            "addc.u32 %2, %10, 0;\n\t"
            "mov.b64 %0, {%3, l1};\n\t"
            "mov.b64 %1, {m0, m1};\n\t"
            "}"
            : "=l"(w.lo), "=l"(w.mid), "=r"(top) : "r"(e0), "r"(e1), "r"(o0), "r"(o1), "r"(h0), "r"(h1), "r"(c1), "r"(c2));
        w.hi = top;
        return w.reduce();
    }
};
#else
typedef acc192w acc192;
#endif

CSG_HD fe root_of_unity(unsigned logn) {  // primitive 2^logn-th root (winterfell StarkField::get_root_of_unity)
    fe r = to_mont(TWO_ADIC_ROOT);
    for (unsigned i = logn; i < TWO_ADICITY; i++) r = sqr(r);
    return r;
}

}  // namespace f63
