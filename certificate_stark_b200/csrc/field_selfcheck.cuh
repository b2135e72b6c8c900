// Self-check of the field arithmetic against its long-hand definition: the fused multiplication (mul, mul_2p) and the
// lazily carried accumulators (acc128, acc192; field.cuh) versus mul_wide + redc_reference + modular additions, on random
// operands and on the patterns that exercise the carry paths (zero / all-ones words, values next to 0, p and 2^64).
// One function for both sides: the device runs it in csg_debug_field_selftest (tests/test_gpu_parity.py), the host-compiled
// copy runs in tests/host_harness.cpp -- which also pins the checker itself, so a device mismatch is a device bug.
#pragma once
#include "field.cuh"

namespace f63 {

CSG_HD uint64_t selfcheck_next(uint64_t &x) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; }
CSG_HD uint64_t selfcheck_operand(uint64_t &x, int pat, bool canonical) {
    uint64_t v = selfcheck_next(x);
    switch (pat & 7) {
    case 1: v <<= 32; break;                    // low word zero
    case 2: v >>= 32; break;                    // high word zero
    case 3: v |= 0xffffffffULL; break;          // low word all ones
    case 4: v = ~(v & 0xffff); break;           // just below 2^64
    case 5: v &= 0xffff; break;                 // just above 0
    case 6: v = P - 1 - (v & 0xffff); break;    // just below p
    default: break;
    }
    return canonical ? v % P : v;
}
CSG_HD fe selfcheck_product(uint64_t a, fe b) {   // a * b * 2^-64 mod p the long way; a any 64-bit value, b < p
    const u128 t = mul_wide(a, b);
    return redc_reference(t.lo, t.hi);
}

// number of mismatches over `iters` rounds
CSG_HD unsigned long long field_selfcheck(uint64_t seed, int iters) {
    uint64_t x = seed | 1;
    unsigned long long bad = 0;
    for (int it = 0; it < iters; it++) {
        const fe a = selfcheck_operand(x, it, true), b = selfcheck_operand(x, it >> 3, true);
        const uint64_t wide = selfcheck_operand(x, it >> 6, false);
        if (mul(a, b) != selfcheck_product(a, b)) bad++;
        if (reduce_2p(mul_2p(wide, b)) != selfcheck_product(wide, b)) bad++;
        // 14 products of canonical operands in acc128, as two halves joined by add()
        acc128 s1, s1b;
        fe r1 = 0;
        for (int k = 0; k < 14; k++) {
            const fe u = selfcheck_operand(x, it + k, true), v = selfcheck_operand(x, (it >> 2) + k, true);
            if (k < 7) s1.mac(u, v); else s1b.mac(u, v);
            r1 = add(r1, selfcheck_product(u, v));
        }
        s1.add(s1b);
        if (s1.reduce() != r1) bad++;
        // 40 products with one arbitrary 64-bit operand in acc192
        acc192 s2;
        fe r2 = 0;
        for (int k = 0; k < 40; k++) {
            const uint64_t u = selfcheck_operand(x, it + 3 * k, false);
            const fe v = selfcheck_operand(x, (it >> 1) + k, true);
            s2.mac(u, v);
            r2 = add(r2, selfcheck_product(u, v));
        }
        if (s2.reduce() != r2) bad++;
    }
    return bad;
}

}  // namespace f63
