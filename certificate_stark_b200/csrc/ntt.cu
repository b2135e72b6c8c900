// See ntt.cuh for the design.  Hand-written for sm_100a: no library FFT exists for this field.
#include <cstdlib>
#include <vector>

#include "ntt.cuh"

namespace csg {
using namespace f63;

namespace {

#ifndef CSG_NTT_TILE_LOG
#define CSG_NTT_TILE_LOG 13
#endif
constexpr unsigned TILE_LOG = CSG_NTT_TILE_LOG;   // elements staged per CTA (2^13 = 64 KB of shared memory + padding)
constexpr unsigned MAX_SUB_LOG = 11;   // largest single sub-transform
constexpr unsigned NTT_THREADS = 256;
#ifndef CSG_NTT_SHAPE_A_DEFAULT
#define CSG_NTT_SHAPE_A_DEFAULT 8
#endif
#ifndef CSG_NTT_SHAPE_B_DEFAULT
#define CSG_NTT_SHAPE_B_DEFAULT 8
#endif

struct PassArgs {
    const fe *in; fe *out;
    unsigned logS, logT;                           // sub-transform size, lanes per CTA
    unsigned nlanes;                               // total lanes along grid.x
    unsigned long long in_se, in_sl, out_se, out_sl;  // element / lane strides, in field elements
    unsigned long long in_by, in_bz, out_by, out_bz;  // strides for blockIdx.y (column) and blockIdx.z (coset)
    const fe *W; unsigned logW;                    // root table
    int inverse;
    const fe *preA, *preB; unsigned long long pre_bz;    // input (e, lane) *= preA[e] * preB[lane]
    const fe *preFull; unsigned long long pre_full_bz;   // or, when given: input at offset m inside its column *= preFull[m]
    const fe *postA, *postB; unsigned long long post_bz;  // output (e, lane) *= postA[e] * postB[lane]
    unsigned tw_logn;                              // != 0: output (e, lane) *= w_{2^tw_logn}^(+-e*lane)
    int use_scalar; fe scalar;                     // output *= scalar
    int clamp_in;                                  // input words are arbitrary u64 (a caller's trace): bring them below 2p first
};

__device__ __forceinline__ fe root_pow(const fe *W, unsigned logW, unsigned logm, unsigned long long e, int inverse) {
    // w_{2^logm}^(+-e), e < 2^logm <= 2^logW
    unsigned long long idx = e << (logW - logm), N = 1ULL << logW;
    if (inverse) idx = (N - idx) & (N - 1);
    return W[idx];
}

// shared-memory index of element i of a lane: one element of padding after every 16 keeps the strided accesses of the
// register rounds (and most of the bit-reversed staging) on distinct banks
__device__ __forceinline__ unsigned pad(unsigned i) { return i + (i >> 4); }
__host__ __device__ inline unsigned lane_pitch(unsigned S) { return S + (S >> 4) + 1; }

// radix-2^R decimation-in-time round: stages s .. s+R-1 of the S-point transform on 2^R elements held in registers.
// Element j of a group sits at base + j*2^s; stage s+t pairs (j, j + 2^t) with twiddle w_S^((k0 + (j mod 2^t)*2^s) * S/2^(s+t+1)).
template <int R>
__device__ __forceinline__ void dit_round(fe *sm, const fe *tw, unsigned logS, unsigned logT, unsigned s, unsigned SP, unsigned tid, unsigned nth) {
    constexpr unsigned E = 1u << R;
    const unsigned groups = 1u << (logS - R + logT), per_lane_mask = (1u << (logS - R)) - 1;
    for (unsigned g = tid; g < groups; g += nth) {
        const unsigned lane = g >> (logS - R), gg = g & per_lane_mask;
        const unsigned k0 = gg & ((1u << s) - 1), base = ((gg >> s) << (s + R)) + k0;
        fe *L = sm + lane * SP;
        fe v[E];
#pragma unroll
        for (unsigned j = 0; j < E; j++) v[j] = L[pad(base + (j << s))];
#pragma unroll
        for (unsigned t = 0; t < (unsigned)R; t++) {
            const unsigned h = 1u << t;
#pragma unroll
            for (unsigned j = 0; j < E; j++) {
                if (j & h) continue;
                const unsigned k = k0 + ((j & (h - 1)) << s);
                const fe w = tw[k << (logS - 1 - s - t)];
                const fe x = v[j], y = mul_2p(v[j + h], w);   // lazy butterfly: values stay in [0, 2p) between the rounds
                v[j] = add_2p(x, y);
                v[j + h] = sub_2p(x, y);
            }
        }
#pragma unroll
        for (unsigned j = 0; j < E; j++) L[pad(base + (j << s))] = v[j];
    }
}

// One pass: T lanes x S elements staged in shared memory, S-point natural-order NTT along each lane (bit-reversed on
// the way in, decimation-in-time rounds of three stages in registers), optional scalings on the way in and out.
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(PassArgs a) {
    extern __shared__ fe sm[];
    const unsigned S = 1u << a.logS, T = 1u << a.logT, SP = lane_pitch(S);
    fe *tw = sm + (size_t)T * SP;                                  // S/2 twiddles of the sub-transform
    const unsigned tid = threadIdx.x, nth = blockDim.x;
    const unsigned lane0 = blockIdx.x << a.logT;
    const fe *in = a.in + blockIdx.y * a.in_by + blockIdx.z * a.in_bz;
    fe *out = a.out + blockIdx.y * a.out_by + blockIdx.z * a.out_bz;

    for (unsigned j = tid; j < S / 2; j += nth) tw[j] = root_pow(a.W, a.logW, a.logS, j, a.inverse);

    const fe *preA = a.preA ? a.preA + blockIdx.z * a.pre_bz : nullptr;
    const fe *preB = a.preB ? a.preB + blockIdx.z * a.pre_bz : nullptr;
    const bool lane_fast_in = a.in_sl == 1 && T > 1;
#pragma unroll 8   // several global loads in flight per thread: the staging phase is latency-bound otherwise
    for (unsigned idx = tid; idx < S * T; idx += nth) {
        unsigned e, l;
        if (lane_fast_in) { l = idx & (T - 1); e = idx >> a.logT; } else { e = idx & (S - 1); l = idx >> a.logS; }
        fe v = 0;
        if (lane0 + l < a.nlanes) {
            const unsigned long long m = e * a.in_se + (lane0 + l) * a.in_sl;
            v = in[m];
            if (a.clamp_in && v >= 2 * P) v -= 2 * P;
            if (a.preFull) v = mul(v, a.preFull[blockIdx.z * a.pre_full_bz + m]);
            else {
                if (preA) v = mul(v, preA[e]);
                if (preB) v = mul(v, preB[lane0 + l]);
            }
        }
        unsigned r = a.logS ? (__brev(e) >> (32 - a.logS)) : 0;
        sm[l * SP + pad(r)] = v;
    }
    __syncthreads();

    // logS = 3*q + rem: the rem (1 or 2) lowest stages go first as one small round, then q rounds of three stages
    unsigned s = 0;
    const unsigned rem = a.logS % 3;
    if (rem == 1) { dit_round<1>(sm, tw, a.logS, a.logT, 0, SP, tid, nth); s = 1; __syncthreads(); }
    else if (rem == 2) { dit_round<2>(sm, tw, a.logS, a.logT, 0, SP, tid, nth); s = 2; __syncthreads(); }
    for (; s < a.logS; s += 3) {
        dit_round<3>(sm, tw, a.logS, a.logT, s, SP, tid, nth);
        __syncthreads();
    }

    const fe *postA = a.postA ? a.postA + blockIdx.z * a.post_bz : nullptr;
    const fe *postB = a.postB ? a.postB + blockIdx.z * a.post_bz : nullptr;
    const bool lane_fast_out = a.out_sl == 1 && T > 1;
    for (unsigned idx = tid; idx < S * T; idx += nth) {
        unsigned e, l;
        if (lane_fast_out) { l = idx & (T - 1); e = idx >> a.logT; } else { e = idx & (S - 1); l = idx >> a.logS; }
        if (lane0 + l >= a.nlanes) continue;
        fe v = sm[l * SP + pad(e)];   // in [0, 2p): a multiplication below brings it under p, otherwise one subtraction does
        if (!(a.tw_logn || postA || postB || a.use_scalar)) v = reduce_2p(v);
        if (a.tw_logn) v = mul(v, root_pow(a.W, a.logW, a.tw_logn, (unsigned long long)e * (lane0 + l), a.inverse));
        if (postA) v = mul(v, postA[e]);
        if (postB) v = mul(v, postB[lane0 + l]);
        if (a.use_scalar) v = mul(v, a.scalar);
        out[e * a.out_se + (lane0 + l) * a.out_sl] = v;
    }
}


// ---------------------------------------------------------------------------------------------- 1024-point passes, one warp per lane
// The two-pass sizes that matter most (n = 2^19 .. 2^21: trace lengths of 512-2048 transactions) have 1024-point
// sub-transforms.  For those a warp owns one lane and does the whole sub-transform as 32 x 32 (four-step inside the warp):
//     i = 32a + b,  k = c + 32d:   X[c + 32d] = sum_b w32^(bd) * ( w1024^(bc) * sum_a w32^(ac) x[32a + b] )
// thread b runs a 32-point transform over a entirely in registers (decimation in frequency, lazy butterflies, the 31 unit
// twiddles skipped), multiplies by w1024^(bc), the warp transposes through its own 8 KB of shared memory (XOR swizzle,
// conflict free, __syncwarp only), thread c runs the second 32-point transform over b and the warp leaves X in natural order
// in its lane of the tile.  Two block-wide barriers and three trips through shared memory per pass instead of five and five;
// 4 modular multiplications per element instead of 5.  Lanes whose elements are contiguous in memory (pass B) are read
// straight from global memory; strided lanes (pass A) are staged as a tile of 8 adjacent lanes, as are all outputs.
struct Fft32Tw { fe w[16]; };   // w32^j, j < 16 (forward or inverse)


// 32-point transform in registers, decimation in time: v[r] = x[brev5(r)] on entry, v[k] = X[k] on exit; values in [0, 2p)
__device__ __forceinline__ void fft32_dit(uint64_t (&v)[32], const Fft32Tw &tw) {
#pragma unroll
    for (int t = 0; t < 5; t++) {
        const int h = 1 << t;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (j & h) continue;
            const int ti = (j & (h - 1)) << (4 - t);
            const uint64_t x = v[j];
            if (ti) {
                const uint64_t y = mul_2p(v[j + h], tw.w[ti]);
                v[j] = add_2p(x, y);
                v[j + h] = sub_2p(x, y);
            } else {   // unit twiddle: no multiplication, so the sum of two values below 2p needs the overflow-safe form
                const uint64_t y = v[j + h];
                v[j] = add_2p_any(x, y);
                v[j + h] = sub_2p(x, y);
            }
        }
    }
}
__device__ __forceinline__ constexpr int brev5(int r) { return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4); }

// F_LANES lanes (= warps) per CTA; lane pitch chosen so that the staged accesses (lane fastest) of a half-warp hit distinct banks
template <int F_LANES> struct FastShape { static constexpr unsigned SP = F_LANES == 8 ? 1024 + 2 : 1024 + 4, THREADS = 32 * F_LANES, LOG = F_LANES == 8 ? 3 : 2; };

// MINB: resident CTAs per SM the register allocation is sized for (2 x 8 lanes and 4 x 4 lanes = 16 warps at 128 registers;
// 5 x 4 lanes = 20 warps at 102 registers, still without spills; shared memory allows no more than 5 CTAs of 4 lanes)
template <bool STAGE_IN, int F_LANES, int MINB>
__global__ void __launch_bounds__(32 * F_LANES, MINB) ntt1024_kernel(PassArgs a, Fft32Tw tw32) {
    constexpr unsigned F_SP = FastShape<F_LANES>::SP, NTH = FastShape<F_LANES>::THREADS, LLOG = FastShape<F_LANES>::LOG;
    extern __shared__ fe sm[];
    fe *T2 = sm + (size_t)F_LANES * F_SP;          // w1024^(+-b c) at [c*32 + b]
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lane0 = blockIdx.x * F_LANES;
    const fe *in = a.in + blockIdx.y * a.in_by + blockIdx.z * a.in_bz;
    fe *out = a.out + blockIdx.y * a.out_by + blockIdx.z * a.out_bz;
    const fe *preA = a.preA ? a.preA + blockIdx.z * a.pre_bz : nullptr;
    const fe *preB = a.preB ? a.preB + blockIdx.z * a.pre_bz : nullptr;

    for (unsigned j = tid; j < 1024; j += NTH) T2[j] = root_pow(a.W, a.logW, 10, ((j >> 5) * (j & 31)) & 1023, a.inverse);

    // The coset pre-scale s^m of element m = e*n2 + lane factors as preA[e] * preB[lane].  A per-lane factor commutes with the
    // transform along the lane, so in a staged pass with an output twiddle (pass A) only preA[e] (a 1024-entry table that
    // stays in L1) is applied on the way in and preB[lane] joins the twiddle on the way out: the data are then the only
    // stream read from HBM (the full per-coset table and the twiddle gather of the generic kernel are not needed).
    const bool fold = a.tw_logn && preA && preB;
    const bool chain_out = a.tw_logn && !a.postA && !a.postB && !a.use_scalar;   // pass A: the output twiddle is a per-thread geometric chain
    fe *PF = T2 + 1024;                                                      // its seeds, [3][NTH]
    auto scaled = [&](fe v, unsigned e, unsigned l) -> fe {   // the input scaling of element e of lane l of this tile
        if (a.clamp_in && v >= 2 * P) v -= 2 * P;
        if (fold) return mul(v, preA[e]);
        if (a.preFull) return mul(v, a.preFull[blockIdx.z * a.pre_full_bz + e * a.in_se + (lane0 + l) * a.in_sl]);
        if (preA) v = mul(v, preA[e]);
        if (preB) v = mul(v, preB[lane0 + l]);
        return v;
    };
    auto load_scaled = [&](unsigned e, unsigned l) -> fe { return scaled(in[e * a.in_se + (lane0 + l) * a.in_sl], e, l); };
    if (STAGE_IN) {
        // The whole tile travels global -> shared as asynchronous 8-byte copies, all of a thread's 32 in flight at once (no
        // registers held): staged through registers 8 at a time, 60 % of this kernel's stall samples sat on the four
        // exposed round trips to HBM (ncu source view).  The scaling moves to the point where a warp picks up its lane.
#pragma unroll
        for (unsigned r = 0; r < 1024 * F_LANES / NTH; r++) {
            const unsigned idx = tid + r * NTH, l = idx & (F_LANES - 1), e = idx >> LLOG;
            const bool ok = lane0 + l < a.nlanes;
            const fe *src = ok ? in + (e * a.in_se + (lane0 + l) * a.in_sl) : in;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + l * F_SP + e);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(ok ? 8u : 0u) : "memory");
        }
        if (chain_out) {
            // the seeds of this thread's output twiddle chain (two roots and the lane's pre-scale) come along: read at the start
            // of the output phase they cost every CTA one more exposed round trip
            const unsigned l = tid & (F_LANES - 1), e0 = tid >> LLOG, L = lane0 + l;
            if (L < a.nlanes) {
                const unsigned long long N = 1ULL << a.logW;
                unsigned long long i_step = (32ULL * L) << (a.logW - a.tw_logn), i_g0 = ((unsigned long long)e0 * L) << (a.logW - a.tw_logn);
                if (a.inverse) { i_step = (N - i_step) & (N - 1); i_g0 = (N - i_g0) & (N - 1); }
                const fe *srcs[3] = {a.W + i_step, a.W + i_g0, fold ? preB + L : a.W};   // W[0] = 1 when there is no pre-scale
#pragma unroll
                for (int k = 0; k < 3; k++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(PF + k * NTH + tid)), "l"(srcs[k]) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    fe *mine = sm + warp * F_SP;
    if (lane0 + warp < a.nlanes) {
        uint64_t v[32];
        // register r: a = brev5(r).  The two common cases first, each a compact run of code: the unrolled general form with its
        // three optional scalings per element spreads the executed instructions over far more instruction-cache lines
        if (!STAGE_IN && !a.preFull && !preA && !preB) {
            const fe *src = in + (lane0 + warp) * a.in_sl;
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = src[(32 * brev5(k) + lane) * a.in_se];
        } else if (STAGE_IN && fold) {
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = mul(mine[32 * brev5(k) + lane], preA[32 * brev5(k) + lane]);
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = STAGE_IN ? scaled(mine[32 * brev5(k) + lane], 32 * brev5(k) + lane, warp) : load_scaled(32 * brev5(k) + lane, warp);
        }
        __syncwarp();
#pragma unroll 1
        for (int phase = 0; phase < 2; phase++) {
            fft32_dit(v, tw32);
            if (phase == 0) {
                // register c holds Y_b[c], b = lane: twiddle, then hand element (b, c) to thread c
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const uint64_t val = c ? mul_2p(v[c], T2[c * 32 + lane]) : v[c];
                    mine[lane * 32 + (c ^ lane)] = val;
                }
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 32; r++) v[r] = mine[brev5(r) * 32 + (lane ^ brev5(r))];   // register r: b = brev5(r)
                __syncwarp();
            }
        }
        // register d holds X[c + 32 d], c = lane
#pragma unroll
        for (int dd = 0; dd < 32; dd++) mine[lane + 32 * dd] = v[dd];
    }
    __syncthreads();

    if (chain_out) {
        // pass A: element e of lane L is multiplied by w_n^(+-e L) (times preB[L] when folded).  A thread keeps one lane and
        // walks e in steps of 32: the factors form a geometric sequence, two interleaved chains, no table gather
        const unsigned l = tid & (F_LANES - 1), e0 = tid >> LLOG, L = lane0 + l;
        if (L < a.nlanes) {
            const fe step = STAGE_IN ? PF[tid] : root_pow(a.W, a.logW, a.tw_logn, 32ULL * L, a.inverse), step2 = sqr(step);
            fe g0 = STAGE_IN ? PF[NTH + tid] : root_pow(a.W, a.logW, a.tw_logn, (unsigned long long)e0 * L, a.inverse);
            if (fold) g0 = mul(g0, STAGE_IN ? PF[2 * NTH + tid] : preB[L]);
            fe g1 = mul(g0, step);
            fe *o = out + L * a.out_sl;
            const fe *src = sm + l * F_SP;
#pragma unroll 4
            for (unsigned k = 0; k < 32; k += 2) {
                const unsigned ea = e0 + 32 * k, eb = ea + 32;
                o[ea * a.out_se] = mul(src[ea], g0);
                o[eb * a.out_se] = mul(src[eb], g1);
                g0 = mul(g0, step2); g1 = mul(g1, step2);
            }
        }
        return;
    }
    const fe *postA = a.postA ? a.postA + blockIdx.z * a.post_bz : nullptr;
    const fe *postB = a.postB ? a.postB + blockIdx.z * a.post_bz : nullptr;
#pragma unroll 4
    for (unsigned idx = tid; idx < 1024 * F_LANES; idx += NTH) {
        const unsigned l = idx & (F_LANES - 1), e = idx >> LLOG;
        if (lane0 + l >= a.nlanes) continue;
        fe v = sm[l * F_SP + e];   // in [0, 2p)
        if (!(a.tw_logn || postA || postB || a.use_scalar)) v = reduce_2p(v);
        if (a.tw_logn) v = mul(v, root_pow(a.W, a.logW, a.tw_logn, (unsigned long long)e * (lane0 + l), a.inverse));
        if (postA) v = mul(v, postA[e]);
        if (postB) v = mul(v, postB[lane0 + l]);
        if (a.use_scalar) v = mul(v, a.scalar);
        out[e * a.out_se + (lane0 + l) * a.out_sl] = v;
    }
}

// the warp-per-lane kernel applies when the sub-transform has 1024 points, adjacent lanes are adjacent in the output
// (tile store), and the input is either a tile of adjacent lanes too (pass A) or contiguous per lane (pass B)
bool fast1024_applies(const PassArgs &a) {
    static const bool off = getenv("CSG_NTT_GENERIC") != nullptr;   // A/B testing against the generic kernel
    if (off || a.logS != 10 || a.out_sl != 1 || a.nlanes < 8) return false;
    return a.in_sl == 1 || a.in_se == 1;
}
template <bool STAGE_IN, int F_LANES, int MINB>
void launch_fast1024_as(const PassArgs &a, const Fft32Tw &tw, unsigned ncols, unsigned ncosets, Stream &st) {
    const size_t smem = ((size_t)F_LANES * FastShape<F_LANES>::SP + 1024 + 3 * FastShape<F_LANES>::THREADS) * sizeof(fe);
    dim3 grid((a.nlanes + F_LANES - 1) / F_LANES, ncols, ncosets);
    CSG_CUDA(cudaFuncSetAttribute(ntt1024_kernel<STAGE_IN, F_LANES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CSG_LAUNCH(st, (ntt1024_kernel<STAGE_IN, F_LANES, MINB>), grid, FastShape<F_LANES>::THREADS, smem, a, tw);
}
void launch_fast1024(const PassArgs &a, unsigned ncols, unsigned ncosets, Stream &st) {
    Fft32Tw tw;
    fe w32 = root_of_unity(5);
    if (a.inverse) w32 = inv(w32);
    fe acc = ONE;
    for (int j = 0; j < 16; j++) { tw.w[j] = acc; acc = mul(acc, w32); }
    // shape of a pass: 8 = two CTAs of 8 lanes, 4 = four CTAs of 4 lanes, 5 = five CTAs of 4 lanes per SM (tuning knobs for A/B runs)
    static const int shape_a = getenv("CSG_NTT_SHAPE_A") ? atoi(getenv("CSG_NTT_SHAPE_A")) : CSG_NTT_SHAPE_A_DEFAULT;
    static const int shape_b = getenv("CSG_NTT_SHAPE_B") ? atoi(getenv("CSG_NTT_SHAPE_B")) : CSG_NTT_SHAPE_B_DEFAULT;
    if (a.in_sl == 1) {
        if (shape_a == 8) launch_fast1024_as<true, 8, 2>(a, tw, ncols, ncosets, st);
        else if (shape_a == 4) launch_fast1024_as<true, 4, 4>(a, tw, ncols, ncosets, st);
        else launch_fast1024_as<true, 4, 5>(a, tw, ncols, ncosets, st);
    } else {
        if (shape_b == 8) launch_fast1024_as<false, 8, 2>(a, tw, ncols, ncosets, st);
        else if (shape_b == 4) launch_fast1024_as<false, 4, 4>(a, tw, ncols, ncosets, st);
        else launch_fast1024_as<false, 4, 5>(a, tw, ncols, ncosets, st);
    }
}

void launch_pass(const PassArgs &a, unsigned ncols, unsigned ncosets, Stream &st) {
    if (fast1024_applies(a)) { launch_fast1024(a, ncols, ncosets, st); return; }
    const unsigned S = 1u << a.logS, T = 1u << a.logT;
    size_t smem = ((size_t)T * lane_pitch(S) + S / 2 + 1) * sizeof(fe);
    if (smem > 48 * 1024)   // per device and cheap: set whenever a launch needs more than the default 48 KB
        CSG_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    unsigned threads = NTT_THREADS;
    while (threads > 32 && threads > (S * T) / 8) threads >>= 1;
    dim3 grid((a.nlanes + T - 1) / T, ncols, ncosets);
    CSG_LAUNCH(st, ntt_pass_kernel, grid, threads, smem, a);
}

// W[j] = base_hi[j >> 10] * base_lo[j & 1023]
__global__ void roots_kernel(fe *W, const fe *lo, const fe *hi, unsigned long long n) {
    unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (j < n) W[j] = mul(hi[j >> 10], lo[j & 1023]);
}

// scale tables for a coset shift s and a split n = n1*n2 (element m = i1*n2 + i2):  A[i1] = s^(i1*n2),  B[i2] = s^i2
__global__ void scale_tables_kernel(const fe *shifts, unsigned n1, unsigned n2, fe *tables) {
    const fe s = shifts[blockIdx.y];
    fe *A = tables + (size_t)blockIdx.y * (n1 + n2), *B = A + n1;
    unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n1) A[j] = f63::pow(s, (unsigned long long)j * n2);
    else if (j < n1 + n2) B[j - n1] = f63::pow(s, j - n1);
}

struct Split { unsigned l1, l2; };
Split split_of(unsigned logn) {
    if (logn <= MAX_SUB_LOG) return {logn, 0};
    unsigned l1 = (logn + 1) / 2;
    if (l1 > MAX_SUB_LOG || logn - l1 > MAX_SUB_LOG) throw std::runtime_error("NTT size above 2^22 is not supported");
    return {l1, logn - l1};
}
unsigned lanes_log(unsigned logS, unsigned nlanes_log) {
    unsigned t = TILE_LOG > logS ? TILE_LOG - logS : 0;
    return t < nlanes_log ? t : nlanes_log;
}
void upload_shifts(NttScratch &sc, const fe *host, size_t n, Stream &st) {
    sc.shifts.reserve(n < 64 ? 64 : n);
    CSG_CUDA(cudaMemcpyAsync(sc.shifts.p, host, n * sizeof(fe), cudaMemcpyHostToDevice, st.s));
}
void build_scale_tables(NttScratch &sc, unsigned n1, unsigned n2, size_t ncosets, Stream &st) {
    sc.scale.reserve((size_t)(n1 + n2) * ncosets);
    dim3 grid((n1 + n2 + 127) / 128, (unsigned)ncosets);
    CSG_LAUNCH(st, scale_tables_kernel, grid, 128, 0, sc.shifts.p, n1, n2, sc.scale.p);
}

}  // namespace

void RootTable::build(unsigned logn_, Stream &st) {
    if (logn_ == logn && W.p) return;
    logn = logn_;
    const size_t n = (size_t)1 << logn;
    W.reserve(n);
    std::vector<fe> lo(1024), hi(((n + 1023) >> 10));
    fe w = root_of_unity(logn), acc = ONE;
    for (size_t i = 0; i < 1024; i++) { lo[i] = acc; acc = mul(acc, w); }
    fe step = acc;  // w^1024
    acc = ONE;
    for (size_t i = 0; i < hi.size(); i++) { hi[i] = acc; acc = mul(acc, step); }
    DBuf<fe> dlo, dhi;
    dlo.reserve(1024); dhi.reserve(hi.size());
    CSG_CUDA(cudaMemcpyAsync(dlo.p, lo.data(), 1024 * sizeof(fe), cudaMemcpyHostToDevice, st.s));
    CSG_CUDA(cudaMemcpyAsync(dhi.p, hi.data(), hi.size() * sizeof(fe), cudaMemcpyHostToDevice, st.s));
    CSG_LAUNCH(st, roots_kernel, (unsigned)((n + 255) / 256), 256, 0, W.p, dlo.p, dhi.p, (unsigned long long)n);
    CSG_CUDA(cudaStreamSynchronize(st.s));  // dlo/dhi/lo/hi go out of scope
}

void intt_columns(const RootTable &rt, NttScratch &sc, const fe *in, size_t in_stride, fe *out, size_t out_stride, size_t ncols,
                  unsigned logn, Stream &st, fe out_factor, bool raw_input) {
    if (logn > rt.logn) throw std::runtime_error("root table too small");
    const size_t n = (size_t)1 << logn;
    // the transform is linear: a constant factor on the input (the representation change of a caller's trace) rides on the 1/n scaling
    const fe ninv = out_factor ? mul(inv(to_mont(n % P)), out_factor) : inv(to_mont(n % P));
    Split sp = split_of(logn);
    PassArgs a{};
    a.W = rt.W.p; a.logW = rt.logn; a.inverse = 1;
    if (sp.l2 == 0) {  // single pass: lanes are columns
        a.in = in; a.out = out; a.logS = logn; a.logT = lanes_log(logn, 31); a.nlanes = (unsigned)ncols;
        a.in_se = 1; a.in_sl = in_stride; a.out_se = 1; a.out_sl = out_stride;
        a.use_scalar = 1; a.scalar = ninv; a.clamp_in = raw_input;
        launch_pass(a, 1, 1, st);
        return;
    }
    const size_t n1 = (size_t)1 << sp.l1, n2 = (size_t)1 << sp.l2;
    sc.tmp.reserve(ncols * n);
    // pass A: sub-transforms over i1 (stride n2) for T adjacent i2, twiddle w_n^-(k1*i2), tmp[k1*n2 + i2]
    a.clamp_in = raw_input;
    a.in = in; a.out = sc.tmp.p; a.logS = sp.l1; a.logT = lanes_log(sp.l1, sp.l2); a.nlanes = (unsigned)n2;
    a.in_se = n2; a.in_sl = 1; a.out_se = n2; a.out_sl = 1; a.in_by = in_stride; a.out_by = n; a.tw_logn = logn;
    launch_pass(a, (unsigned)ncols, 1, st);
    // pass B: sub-transforms over i2 (contiguous) for T adjacent k1, out[k2*n1 + k1]
    PassArgs b{};
    b.W = rt.W.p; b.logW = rt.logn; b.inverse = 1;
    b.in = sc.tmp.p; b.out = out; b.logS = sp.l2; b.logT = lanes_log(sp.l2, sp.l1); b.nlanes = (unsigned)n1;
    b.in_se = 1; b.in_sl = n2; b.out_se = n1; b.out_sl = 1; b.in_by = n; b.out_by = out_stride;
    b.use_scalar = 1; b.scalar = ninv;
    launch_pass(b, (unsigned)ncols, 1, st);
}

// full[z][m] = A_z[m / n2] * B_z[m % n2] = shift_z^m
__global__ void full_scale_kernel(const fe *tables, unsigned n1, unsigned n2, fe *full) {
    const unsigned long long m = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x, n = (unsigned long long)n1 * n2;
    if (m >= n) return;
    const fe *A = tables + (size_t)blockIdx.y * (n1 + n2), *B = A + n1;
    full[blockIdx.y * n + m] = mul(A[m / n2], B[m % n2]);
}

void CosetTables::build(const fe *shifts_host, size_t ncosets_, unsigned logn_, Stream &st) {
    Split sp = split_of(logn_);
    const unsigned n1 = 1u << sp.l1, n2 = 1u << sp.l2;
    ncosets = ncosets_; logn = logn_;
    shifts.reserve(ncosets < 64 ? 64 : ncosets);
    tables.reserve((size_t)(n1 + n2) * ncosets);
    CSG_CUDA(cudaMemcpyAsync(shifts.p, shifts_host, ncosets * sizeof(fe), cudaMemcpyHostToDevice, st.s));
    dim3 grid((n1 + n2 + 127) / 128, (unsigned)ncosets);
    CSG_LAUNCH(st, scale_tables_kernel, grid, 128, 0, shifts.p, n1, n2, tables.p);
    has_full = sp.l2 != 0;   // two-pass sizes: one table lookup and one multiplication per element instead of two
    if (has_full) {
        const size_t n = (size_t)n1 * n2;
        full.reserve(n * ncosets);
        CSG_LAUNCH(st, full_scale_kernel, dim3((unsigned)((n + 255) / 256), (unsigned)ncosets), 256, 0, (const fe *)tables.p, n1, n2, full.p);
    }
}

void coset_ntt_columns(const RootTable &rt, NttScratch &sc, const fe *coeffs, size_t in_stride, fe *out, size_t out_col_stride,
                       size_t out_coset_stride, size_t ncols, unsigned logn, const fe *shifts_host, size_t ncosets, Stream &st) {
    // the convenience form builds the tables into the scratch (they are a few KB)
    Split sp = split_of(logn);
    upload_shifts(sc, shifts_host, ncosets, st);
    build_scale_tables(sc, 1u << sp.l1, 1u << sp.l2, ncosets, st);
    coset_ntt_columns(rt, sc, coeffs, in_stride, out, out_col_stride, out_coset_stride, ncols, logn, sc.scale.p, ncosets, st, 0);
}

void coset_ntt_columns(const RootTable &rt, NttScratch &sc, const fe *coeffs, size_t in_stride, fe *out, size_t out_col_stride,
                       size_t out_coset_stride, size_t ncols, unsigned logn, const CosetTables &ct, Stream &st) {
    if (ct.logn != logn) throw std::runtime_error("coset tables were built for another size");
    coset_ntt_columns(rt, sc, coeffs, in_stride, out, out_col_stride, out_coset_stride, ncols, logn, ct.tables.p, ct.ncosets, st, 0,
                      ct.has_full ? ct.full.p : nullptr);
}

void coset_ntt_entries(const RootTable &rt, NttScratch &sc, const fe *in, fe *out, size_t nentries, unsigned logn, const fe *shifts_host, Stream &st) {
    Split sp = split_of(logn);
    const size_t n = (size_t)1 << logn;
    upload_shifts(sc, shifts_host, nentries, st);
    build_scale_tables(sc, 1u << sp.l1, 1u << sp.l2, nentries, st);
    coset_ntt_columns(rt, sc, in, n, out, n, n, 1, logn, sc.scale.p, nentries, st, (int)n);
}

void coset_ntt_columns(const RootTable &rt, NttScratch &sc, const fe *coeffs, size_t in_stride, fe *out, size_t out_col_stride,
                       size_t out_coset_stride, size_t ncols, unsigned logn, const fe *tables_dev, size_t ncosets, Stream &st, int in_coset_stride,
                       const fe *full_tables) {
    if (logn > rt.logn) throw std::runtime_error("root table too small");
    const size_t n = (size_t)1 << logn;
    Split sp = split_of(logn);
    const unsigned n1 = 1u << sp.l1, n2 = 1u << sp.l2;
    PassArgs a{};
    a.W = rt.W.p; a.logW = rt.logn; a.inverse = 0;
    a.preA = tables_dev; a.pre_bz = n1 + n2;
    if (sp.l2 == 0) {
        a.in = coeffs; a.out = out; a.logS = logn; a.logT = lanes_log(logn, 31); a.nlanes = (unsigned)ncols;
        a.in_se = 1; a.in_sl = in_stride; a.out_se = 1; a.out_sl = out_col_stride; a.out_bz = out_coset_stride; a.in_bz = (unsigned long long)in_coset_stride;
        launch_pass(a, 1, (unsigned)ncosets, st);
        return;
    }
    // cosets are processed in groups so that the pass-A scratch stays bounded (4 GB of elements at most)
    size_t group = ncosets;
    const size_t budget = (size_t)1 << 29;
    while (group > 1 && group * ncols * n > budget) group = (group + 1) / 2;
    sc.tmp.reserve(group * ncols * n);
    for (size_t z0 = 0; z0 < ncosets; z0 += group) {
        const size_t g = z0 + group <= ncosets ? group : ncosets - z0;
        a.preA = tables_dev + z0 * (n1 + n2); a.preB = a.preA + n1;
        a.preFull = full_tables ? full_tables + z0 * n : nullptr; a.pre_full_bz = n;
        a.in = coeffs + z0 * (size_t)in_coset_stride; a.out = sc.tmp.p; a.logS = sp.l1; a.logT = lanes_log(sp.l1, sp.l2); a.nlanes = n2;
        a.in_se = n2; a.in_sl = 1; a.out_se = n2; a.out_sl = 1; a.in_by = in_stride; a.in_bz = (unsigned long long)in_coset_stride; a.out_by = n; a.out_bz = ncols * n;
        a.tw_logn = logn;
        launch_pass(a, (unsigned)ncols, (unsigned)g, st);
        PassArgs b{};
        b.W = rt.W.p; b.logW = rt.logn; b.inverse = 0;
        b.in = sc.tmp.p; b.out = out + z0 * out_coset_stride; b.logS = sp.l2; b.logT = lanes_log(sp.l2, sp.l1); b.nlanes = n1;
        b.in_se = 1; b.in_sl = n2; b.out_se = n1; b.out_sl = 1; b.in_by = n; b.in_bz = ncols * n; b.out_by = out_col_stride; b.out_bz = out_coset_stride;
        launch_pass(b, (unsigned)ncols, (unsigned)g, st);
    }
}

void coset_intt_columns(const RootTable &rt, NttScratch &sc, const fe *in, size_t in_stride, fe *out, size_t out_stride,
                        unsigned logn, const fe *shift_inv_host, size_t ncosets, Stream &st) {
    if (logn > rt.logn) throw std::runtime_error("root table too small");
    const size_t n = (size_t)1 << logn;
    const fe ninv = inv(to_mont(n % P));
    Split sp = split_of(logn);
    const unsigned n1 = 1u << sp.l1, n2 = 1u << sp.l2;
    upload_shifts(sc, shift_inv_host, ncosets, st);
    PassArgs a{};
    a.W = rt.W.p; a.logW = rt.logn; a.inverse = 1;
    if (sp.l2 == 0) {
        // output element m *= shift_inv^m: tables A[m] (n1 = n entries), B unused
        build_scale_tables(sc, n1, 1, ncosets, st);
        a.in = in; a.out = out; a.logS = logn; a.logT = 0; a.nlanes = 1;
        a.in_se = 1; a.in_sl = 0; a.out_se = 1; a.out_sl = 0; a.in_bz = in_stride; a.out_bz = out_stride;
        a.postA = sc.scale.p; a.post_bz = n1 + 1;
        a.use_scalar = 1; a.scalar = ninv;
        launch_pass(a, 1, (unsigned)ncosets, st);
        return;
    }
    // output index m = k1 + n1*k2: pass B has e = k2, lane = k1, so it needs  postA[k2] = s^(n1*k2), postB[k1] = s^k1:
    // the tables of the split (n2, n1)
    build_scale_tables(sc, n2, n1, ncosets, st);
    sc.tmp.reserve(ncosets * n);
    a.in = in; a.out = sc.tmp.p; a.logS = sp.l1; a.logT = lanes_log(sp.l1, sp.l2); a.nlanes = n2;
    a.in_se = n2; a.in_sl = 1; a.out_se = n2; a.out_sl = 1; a.in_bz = in_stride; a.out_bz = n; a.tw_logn = logn;
    launch_pass(a, 1, (unsigned)ncosets, st);
    PassArgs b{};
    b.W = rt.W.p; b.logW = rt.logn; b.inverse = 1;
    b.in = sc.tmp.p; b.out = out; b.logS = sp.l2; b.logT = lanes_log(sp.l2, sp.l1); b.nlanes = n1;
    b.in_se = 1; b.in_sl = n2; b.out_se = n1; b.out_sl = 1; b.in_bz = n; b.out_bz = out_stride;
    b.postA = sc.scale.p; b.postB = sc.scale.p + n2; b.post_bz = n1 + n2;
    b.use_scalar = 1; b.scalar = ninv;
    launch_pass(b, 1, (unsigned)ncosets, st);
}

}  // namespace csg
