// Integer-pipe microbenchmark for sm_100a (B200): issue rates of the instruction classes the f63 kernels are made of, alone and
// mixed, and of whole NTT butterflies in several formulations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o int_pipes int_pipes.cu && ./int_pipes
//   ncu --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,\
//       smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,gpu__time_duration.sum ./int_pipes
// Each kernel: 148 CTAs x 1024 threads (8 warps per SM sub-partition), ITER trips of a body of 8 independent chains per thread.
// The program prints time only; per-opcode counts of each body come from `cuobjdump -sass` (tools/microbench/README).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../certificate_stark_b200/csrc/field.cuh"
using namespace f63;

#define ITER 1024
#define CHK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t add64_fma(uint64_t x, uint64_t y, uint32_t one) {
    // 64-bit add on the FMA pipe: low word through a wide multiply-add by a runtime 1, high word through a 32-bit multiply-add
    uint64_t s = (uint64_t)(uint32_t)y * one + x;
    uint32_t hi = (uint32_t)(y >> 32) * one + (uint32_t)(s >> 32);
    return ((uint64_t)hi << 32) | (uint32_t)s;
}
// butterfly formulations: (x, y, w) -> (x + w y, x - w y); x, y lazily reduced
template <int V> __device__ __forceinline__ void bfly(uint64_t &x, uint64_t &y, uint64_t w, uint32_t one) {
    if (V == 0) {          // current: product below 1.52p, sums brought back below 2p by compare + select
        const uint64_t t = mul_2p(y, w), a = x;
        x = add_2p(a, t); y = sub_2p(a, t);
    } else if (V == 1) {   // product fully reduced, invariant x < 2^63: the fix is "subtract p when bit 63 is set", done with multiply-adds
        const uint64_t t0 = mul_raw(y, w), t = t0 >= P ? t0 - P : t0, nt = P - t, a = x;
        uint64_t s = a + t, d = a + nt;
        uint32_t qs = (uint32_t)(s >> 63), qd = (uint32_t)(d >> 63);
        s = (uint64_t)qs * 0xFFFFFFFFu + s; s = (((uint64_t)((uint32_t)(s >> 32) + qs * 0xBE7FFFFFu)) << 32) | (uint32_t)s;
        d = (uint64_t)qd * 0xFFFFFFFFu + d; d = (((uint64_t)((uint32_t)(d >> 32) + qd * 0xBE7FFFFFu)) << 32) | (uint32_t)d;
        x = s; y = d;
    } else if (V == 2) {   // as 0 with the two 64-bit additions on the FMA pipe
        const uint64_t t = mul_2p(y, w), a = x;
        uint64_t s = add64_fma(a, t, one);
        x = s >= 2 * P ? s - 2 * P : s;
        y = sub_2p(a, t);
    } else if (V == 3) {   // mask form of the conditional corrections (shift + and instead of compare + select).  NOT EXACT: the sign test
                           // misreads a + t < 2p - 2^63 (e.g. both zero); timed only to see what three instructions fewer would buy (8 %)
        const uint64_t t = mul_2p(y, w), a = x;
        uint64_t s = a + t - 2 * P;                                  // a + t < 3.52p: s "negative" exactly when a + t < 2p
        uint64_t m = (uint64_t)((int64_t)s >> 63);                    // wrong when s >= 2^63 legitimately: s < 1.52p < 2^63, fine
        x = s + (m & (2 * P));
        uint64_t d = a - t;
        uint64_t borrow = (uint64_t)0 - (uint64_t)(a < t);
        y = d + (borrow & (2 * P));
    }
}
template <int V> __global__ void __launch_bounds__(1024, 1) kb(uint64_t *out, uint64_t seed, uint32_t one) {
    uint64_t x[8], y[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = (seed + threadIdx.x * 16 + i) % P; y[i] = (seed * 3 + i + threadIdx.x) % P; w[i] = (seed * 5 + 7 * i + 1) % P; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) bfly<V>(x[i], y[i], w[i], one);
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i] ^ y[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

template <int KIND> __global__ void __launch_bounds__(1024, 1) k(uint64_t *out, uint32_t seed, uint32_t one) {
    uint32_t a[8], b[8];
    uint64_t q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 8 + i; b[i] = seed * 3 + i; q[i] = ((uint64_t)a[i] << 32) | b[i]; }
    const uint32_t c = seed | 1, d = seed * 7 + 1;
    const uint64_t cd = ((uint64_t)c << 32) | d;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) q[i] = (uint64_t)(uint32_t)q[i] * c + q[i];                      // IMAD.WIDE.U32 with a 64-bit addend
            else if (KIND == 1) a[i] = a[i] * c + b[i];                                       // IMAD
            else if (KIND == 2) { a[i] = __umulhi(a[i], c) + b[i]; }                          // IMAD.HI
            else if (KIND == 3) { a[i] += c; b[i] += d; }                                     // IADD3 x2
            else if (KIND == 4) { a[i] = (a[i] ^ c) & (a[i] | d); b[i] = (b[i] & c) ^ (b[i] | d); }   // LOP3 x2
            else if (KIND == 5) { a[i] = a[i] >= b[i] ? d : a[i]; b[i] += one; }              // ISETP + SEL (+1 add)
            else if (KIND == 6) q[i] += cd;                                                   // 64-bit add: IADD3 + IADD3.X / IMAD.X
            else if (KIND == 7) q[i] = add64_fma(q[i], cd, one);                              // 64-bit add on the FMA pipe
            else if (KIND == 8) { q[i] = (uint64_t)(uint32_t)q[i] * c + q[i]; a[i] += c; }    // IMAD.WIDE + IADD3
            else if (KIND == 9) { q[i] = (uint64_t)(uint32_t)q[i] * c + q[i]; a[i] += c; b[i] += d; }   // IMAD.WIDE + 2 IADD3
            else if (KIND == 10) { a[i] = a[i] * c + b[i]; b[i] += d; }                       // IMAD + IADD3
            else if (KIND == 11) { a[i] = __funnelshift_l(a[i], b[i], 7); b[i] = __funnelshift_l(b[i], a[i], 9); }   // SHF x2
            else if (KIND == 12) q[i] = mul_raw(q[i], cd >> 2);                               // modular multiplication (lazy), dependent chain
            else if (KIND == 13) q[i] = mul(q[i], cd >> 2);                                   // modular multiplication, reduced
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i] ^ b[i] ^ q[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

// lazily reduced multiply-accumulate chains (field.cuh acc128): NACC independent accumulators per thread, each fed ITER terms
template <int NACC> __global__ void __launch_bounds__(1024, 1) kmac(uint64_t *out, uint64_t seed) {
    acc128 acc[NACC];
    uint64_t a[NACC], b = (seed * 7 + threadIdx.x) % P;
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = (seed + threadIdx.x * 16 + i) % P;
    for (int it = 0; it < ITER / 8; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i].mac(a[i], b);
            b += 0x9e3779b97f4a7c15ULL;
            b &= 0x3fffffffffffffffULL;
        }
#pragma unroll
        for (int i = 0; i < NACC; i++) { a[i] = acc[i].reduce(); acc[i] = acc128(); }   // 8 terms, then one reduction (the MDS rows have 14)
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= a[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}
// the butterfly kernels again with fewer resident warps: THREADS per CTA, one CTA per SM
template <int V, int THREADS> __global__ void __launch_bounds__(THREADS, 1) kbo(uint64_t *out, uint64_t seed, uint32_t one) {
    uint64_t x[8], y[8], w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = (seed + threadIdx.x * 16 + i) % P; y[i] = (seed * 3 + i + threadIdx.x) % P; w[i] = (seed * 5 + 7 * i + 1) % P; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) bfly<V>(x[i], y[i], w[i], one);
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i] ^ y[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

// FP64 pipe: is it a third pipe next to fmaheavy (IMAD) and alu?  DFMA alone, DFMA interleaved with IMAD.WIDE, and the int <-> double conversions
template <int KIND> __global__ void __launch_bounds__(1024, 1) kd(uint64_t *out, uint32_t seed, double dscale) {
    double acc[8], x = 1.0 + seed * 1e-9, y = dscale;
    uint64_t q[8];
    uint32_t c = seed | 1;
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = threadIdx.x + i; q[i] = ((uint64_t)(seed + i) << 32) | threadIdx.x; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) acc[i] = fma(acc[i], x, y);                                                   // DFMA
            else if (KIND == 1) { acc[i] = fma(acc[i], x, y); q[i] = (uint64_t)(uint32_t)q[i] * c + q[i]; }   // DFMA + IMAD.WIDE 1:1
            else if (KIND == 2) { acc[i] = fma(acc[i], x, y); acc[i] = fma(acc[i], y, x); q[i] = (uint64_t)(uint32_t)q[i] * c + q[i]; }   // 2:1
            else if (KIND == 3) { acc[i] = (double)(uint32_t)(q[i] >> 11) * x + acc[i]; q[i] += c; }      // I2F.F64.U32 + DFMA
            else if (KIND == 4) { q[i] += (uint64_t)(long long)(acc[i]); acc[i] = fma(acc[i], x, y); }    // F2I.S64.F64 + DFMA
            else if (KIND == 5) { acc[i] = fma(acc[i], x, y); q[i] = ((uint32_t)q[i] + c) ^ (uint32_t)(q[i] >> 32); }   // DFMA + ALU
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= (uint64_t)__double_as_longlong(acc[i]) ^ q[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

template <class F> int timed(const char *name, F launch) {
    cudaEvent_t a, b;
    CHK(cudaEventCreate(&a)); CHK(cudaEventCreate(&b));
    launch(0);
    CHK(cudaDeviceSynchronize());
    CHK(cudaEventRecord(a));
    for (int r = 0; r < 4; r++) launch(r + 1);
    CHK(cudaEventRecord(b));
    CHK(cudaEventSynchronize(b));
    float ms;
    CHK(cudaEventElapsedTime(&ms, a, b));
    printf("%-14s %9.4f ms per launch   (%d trips x 8 chains x 8 warps per SMSP)\n", name, ms / 4, ITER);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CHK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CHK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("%s, %d SMs, clock attribute %.0f MHz\n", p.name, p.multiProcessorCount, khz / 1e3);
    uint64_t *d;
    CHK(cudaMalloc(&d, 8192));
    const int s = p.multiProcessorCount;
#define RUNK(K) timed("class " #K, [&](int r) { k<K><<<s, 1024>>>(d, 12345 + r, 1); });
    RUNK(0) RUNK(1) RUNK(2) RUNK(3) RUNK(4) RUNK(5) RUNK(6) RUNK(7) RUNK(8) RUNK(9) RUNK(10) RUNK(11) RUNK(12) RUNK(13)
#define RUNB(V) timed("butterfly " #V, [&](int r) { kb<V><<<s, 1024>>>(d, 12345 + r, 1); });
    RUNB(0) RUNB(1) RUNB(2) RUNB(3)
    timed("dfma", [&](int r) { kd<0><<<s, 1024>>>(d, 12345 + r, 0.5); });
    timed("dfma+wide", [&](int r) { kd<1><<<s, 1024>>>(d, 12345 + r, 0.5); });
    timed("2dfma+wide", [&](int r) { kd<2><<<s, 1024>>>(d, 12345 + r, 0.5); });
    timed("i2f+dfma", [&](int r) { kd<3><<<s, 1024>>>(d, 12345 + r, 0.5); });
    timed("f2i+dfma", [&](int r) { kd<4><<<s, 1024>>>(d, 12345 + r, 0.5); });
    timed("dfma+alu", [&](int r) { kd<5><<<s, 1024>>>(d, 12345 + r, 0.5); });
    timed("mac x2", [&](int r) { kmac<2><<<s, 1024>>>(d, 12345 + r); });
    timed("mac x4", [&](int r) { kmac<4><<<s, 1024>>>(d, 12345 + r); });
    timed("mac x8", [&](int r) { kmac<8><<<s, 1024>>>(d, 12345 + r); });
    timed("bfly0 w4", [&](int r) { kbo<0, 512><<<s, 512>>>(d, 12345 + r, 1); });
    timed("bfly0 w2", [&](int r) { kbo<0, 256><<<s, 256>>>(d, 12345 + r, 1); });
    timed("bfly0 w1", [&](int r) { kbo<0, 128><<<s, 128>>>(d, 12345 + r, 1); });
    timed("bfly3 w4", [&](int r) { kbo<3, 512><<<s, 512>>>(d, 12345 + r, 1); });
    CHK(cudaDeviceSynchronize());
    return 0;
}
