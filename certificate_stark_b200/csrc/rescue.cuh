// Rescue-XLIX over f63 (width 14, rate 7, alpha 3, 7 rounds) -- the hash the reference's AIRs arithmetise
// (reference: src/utils/rescue.rs).  Shared by the host witness builder and the constraint kernels.
//
// Device notes: the 14x14 MDS products are accumulated unreduced in 128 bits (14 * p^2 < 2^128) and reduced once
// per output row, so a half round costs 196 wide multiply-adds + 14 reductions instead of 196 full modmuls.
#pragma once
#include "field.cuh"
#include "ref_constants.h"

namespace rescue {
using f63::fe;

constexpr int STATE_WIDTH = 14, RATE_WIDTH = 7, NUM_ROUNDS = 7, CYCLE = 8;

// out = M * in, M = MDS or INV_MDS (row-major, Montgomery form).  On the device the tables live in __constant__ memory
// and the row loop is kept rolled: the 14-term dot product is the unit the instruction cache sees.
template <bool INVERSE>
CSG_HD void mat_mul(const fe (&in)[14], fe (&out)[14]) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1   // fully unrolled (196 multiply-adds) it is 3x the code and spills under the kernels' register caps
#endif
    for (int i = 0; i < 14; i++) {
        const uint64_t *row = (INVERSE ? CSG_TABLE(CSG_INV_MDS) : CSG_TABLE(CSG_MDS)) + i * 14;
        f63::acc128 acc;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 14; j++) acc.mac(row[j], in[j]);
        out[i] = acc.reduce();
    }
}
CSG_HD fe cube(fe x) { return f63::mul(x, f63::sqr(x)); }

// forward half of a round: MDS * sbox(cur) + ark[0..14)        (rescue.rs:274-279)
CSG_HD void forward_half(const fe (&cur)[14], const fe *ark, fe (&out)[14]) {
    fe t[14];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 14; i++) t[i] = cube(cur[i]);
    mat_mul<false>(t, out);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 14; i++) out[i] = f63::add(out[i], ark[i]);
}
// inverse of the second half of a round applied to next: sbox(INV_MDS * (next - ark[14..28)))   (rescue.rs:281-287)
CSG_HD void backward_half(const fe (&next)[14], const fe *ark, fe (&out)[14]) {
    fe t[14], u[14];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 14; i++) t[i] = f63::sub(next[i], ark[14 + i]);
    mat_mul<true>(t, u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 14; i++) out[i] = cube(u[i]);
}
// d[i] = backward_half(next)[i] - forward_half(cur)[i]: zero on every valid round transition (rescue.rs:269-300)
CSG_HD void round_residual(const fe (&cur)[14], const fe (&next)[14], const fe *ark, fe (&d)[14]) {
    fe a[14], b[14];
    forward_half(cur, ark, a);
    backward_half(next, ark, b);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 14; i++) d[i] = f63::sub(b[i], a[i]);
}

// ---- the permutation itself, for witness generation on the host and on the device (rescue.rs:239-263, 108-152)
CSG_HD void apply_round(fe *state, size_t step) {
    const uint64_t *ark = CSG_TABLE(CSG_ARK) + (step % CYCLE) * 28;
    fe s[14], t[14];
    for (int i = 0; i < 14; i++) s[i] = state[i];
    forward_half(s, ark, t);
    for (int i = 0; i < 14; i++) t[i] = f63::pow(t[i], CSG_INV_ALPHA);
    mat_mul<false>(t, s);
    for (int i = 0; i < 14; i++) state[i] = f63::add(s[i], ark[14 + i]);
}
CSG_HD void apply_permutation(fe *state) { for (int i = 0; i < NUM_ROUNDS; i++) apply_round(state, i); }
CSG_HD void merge(const fe *a, const fe *b, fe *out) {
    fe st[14];
    for (int i = 0; i < 7; i++) { st[i] = a[i]; st[7 + i] = b[i]; }
    apply_permutation(st);
    for (int i = 0; i < 7; i++) out[i] = st[i];
}
CSG_HD void digest(const fe *data, size_t n, fe *out) {
    fe st[14] = {0};
    size_t i = 0;
    for (size_t k = 0; k < n; k++) {
        st[i] = f63::add(st[i], data[k]);
        if (++i % RATE_WIDTH == 0) { apply_permutation(st); i = 0; }
    }
    if (i > 0) apply_permutation(st);
    for (int k = 0; k < 7; k++) out[k] = st[k];
}
}  // namespace rescue
