// TransactionMetadata::build_random on the GPU (/root/reference/src/lib.rs:235-464): the account tree
// (a depth-15 Rescue Merkle tree, src/lib.rs:261, leaves = Rescue63::merge of the two halves of an account, :283-290), its update
// by every transfer (:370-422), the authentication paths handed to the witness (:369, :422), and the Schnorr signatures
// (src/schnorr/mod.rs:197-216 with hash_message :247-288).
//
// The host builder does this as ~32 dependent Rescue permutations per transfer (two leaf updates of 15 levels each): 2.9 s for 1024
// transfers.  Here the HISTORY of the tree is computed level by level instead: the host plans, without hashing anything, which node
// versions exist (one per level per update, plus the nodes touched by the initial accounts) and which two child versions each one
// merges (host/batch_plan.hpp); a level is then one launch with one thread per version -- thousands of independent permutations --
// and the paths are gathers of version ids.  16 launches for the tree, one for the signatures (fixed-base windowed r.G from a
// table built once per context), one that writes the packed records witness_gen.cu consumes.  Nothing but the two roots returns
// to the host.  Results are bit-identical to the host builder (tests/test_gpu_batch.py).
#include "ecc.cuh"
#include "rescue.cuh"
#include "witness.cuh"

namespace csg {
using namespace f63;

namespace {

__device__ __forceinline__ const fe *version(const fe *hashes, const fe *defaults, int ref) {
    return ref >= 0 ? hashes + (size_t)ref * 7 : defaults + (size_t)(-ref - 1) * 7;
}

// hash of the empty subtree of every level: level 0 = the all-zero digest, level l+1 = merge(level l, level l)
__global__ void batch_defaults_kernel(unsigned depth, fe *defaults) {
    fe h[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 7; i++) defaults[i] = 0;
    for (unsigned l = 0; l < depth; l++) {
        fe o[7];
        rescue::merge(h, h, o);
        for (int i = 0; i < 7; i++) { h[i] = o[i]; defaults[(size_t)(l + 1) * 7 + i] = o[i]; }
    }
}
__global__ void __launch_bounds__(64) batch_leaf_kernel(const uint64_t *__restrict__ accounts, unsigned n, fe *__restrict__ hashes) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    fe o[7];
    rescue::merge(accounts + (size_t)g * 14, accounts + (size_t)g * 14 + 7, o);   // src/lib.rs:283-290
    for (int i = 0; i < 7; i++) hashes[(size_t)g * 7 + i] = o[i];
}
__global__ void __launch_bounds__(64) batch_merge_kernel(unsigned first, unsigned count, const int *__restrict__ left, const int *__restrict__ right,
                                                         fe *hashes, const fe *__restrict__ defaults) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    const unsigned id = first + g;
    fe o[7];
    rescue::merge(version(hashes, defaults, left[id]), version(hashes, defaults, right[id]), o);
    for (int i = 0; i < 7; i++) hashes[(size_t)id * 7 + i] = o[i];
}

// ---- fixed-base table for r.G: T[j][d] = d * 16^j * G in affine coordinates, j < 64, 1 <= d < 16
__global__ void __launch_bounds__(64) batch_gtable_kernel(fe *__restrict__ table) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 64 * 16) return;
    const unsigned j = g >> 4, d = g & 15;
    fe *out = table + (size_t)g * 12;
    if (!d) { for (int i = 0; i < 12; i++) out[i] = 0; return; }
    const uint64_t *G = CSG_TABLE(CSG_GENERATOR);
    const ecc::fp6 gx = ecc::load6(G), gy = ecc::load6(G + 6);
    ecc::point acc{};
    acc.y.c[0] = ONE;
    for (int b = 3; b >= 0; b--) { acc = ecc::double_point(acc); if ((d >> b) & 1) acc = ecc::add_mixed(acc, gx, gy); }
    for (unsigned k = 0; k < 4 * j; k++) acc = ecc::double_point(acc);
    const ecc::fp6 zi = ecc::inv(acc.z), x = ecc::mul(acc.x, zi), y = ecc::mul(acc.y, zi);
    for (int i = 0; i < 6; i++) { out[i] = x.c[i]; out[6 + i] = y.c[i]; }
}

struct SplitMix64 {
    uint64_t s;
    __device__ uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
};

// one signature per thread (src/schnorr/mod.rs:197-216 as the host builder's sign(): secret key sk in {1,2,3}, r drawn in
// [2^254 + 2^253, 2^255) until s = r - sk*h is a non-negative integer -- the circuit's double-and-add never reduces modulo the group
// order).  out: rx[6] s[4] h[4]
__global__ void __launch_bounds__(32) batch_sign_kernel(const uint64_t *__restrict__ tx_words, unsigned ntx, const fe *__restrict__ gtable, uint64_t *__restrict__ out) {
    const unsigned tx = blockIdx.x * blockDim.x + threadIdx.x;
    if (tx >= ntx) return;
    const uint64_t *W = tx_words + (size_t)tx * 33;
    const unsigned sk = (unsigned)W[31];
    SplitMix64 rng{W[32]};
    fe msg[28];   // sender key | receiver key | delta | nonce | 0 | 0   (src/lib.rs:467-481)
    for (int i = 0; i < 12; i++) { msg[i] = W[i]; msg[12 + i] = W[14 + i]; }
    msg[24] = W[28]; msg[25] = W[13]; msg[26] = 0; msg[27] = 0;
    for (;;) {
        uint64_t r[4] = {rng.next(), rng.next(), rng.next(), rng.next()};
        r[3] = (r[3] & 0x7fffffffffffffffULL) | 0x6000000000000000ULL;
        ecc::point acc{};
        acc.y.c[0] = ONE;
        for (unsigned j = 0; j < 64; j++) {
            const unsigned d = (unsigned)(r[j >> 4] >> (4 * (j & 15))) & 15;
            if (d) { const fe *t = gtable + (size_t)(j * 16 + d) * 12; acc = ecc::add_mixed(acc, ecc::load6(t), ecc::load6(t + 6)); }
        }
        const ecc::fp6 x = ecc::mul(acc.x, ecc::inv(acc.z));
        fe h[7], t[7];
        rescue::digest(x.c, 6, h);                                               // src/schnorr/mod.rs:247-288
        for (int k = 0; k < 4; k++) { rescue::merge(h, msg + 7 * k, t); for (int i = 0; i < 7; i++) h[i] = t[i]; }
        uint64_t hw[4], kh[5];
        for (int i = 0; i < 4; i++) hw[i] = from_mont(h[i]);                     // the 256-bit string of src/schnorr/trace.rs:136-139
        uint64_t carry = 0;
        for (int i = 0; i < 4; i++) { const uint64_t lo = hw[i] * sk, hi = __umul64hi(hw[i], sk), s = lo + carry; kh[i] = s; carry = hi + (s < lo); }
        kh[4] = carry;
        uint64_t s[4], borrow = 0;
        for (int i = 0; i < 4; i++) { const uint64_t a = r[i] - kh[i], b = a - borrow; borrow = (r[i] < kh[i]) | (a < borrow); s[i] = b; }
        if (kh[4] == 0 && borrow == 0) {
            uint64_t *o = out + (size_t)tx * 14;
            for (int i = 0; i < 6; i++) o[i] = x.c[i];
            for (int i = 0; i < 4; i++) { o[6 + i] = s[i]; o[10 + i] = hw[i]; }
            return;
        }
    }
}

// the packed record of every transfer (witness.cuh): thread = (transfer, word)
__global__ void batch_pack_kernel(const uint64_t *__restrict__ tx_words, const int *__restrict__ tx_refs, unsigned ntx, const fe *__restrict__ hashes,
                                  const fe *__restrict__ defaults, const uint64_t *__restrict__ sigs, uint64_t *__restrict__ records) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ntx * WIT_WORDS) return;
    const unsigned tx = g / WIT_WORDS, w = g % WIT_WORDS;
    const uint64_t *W = tx_words + (size_t)tx * 33;
    const int *R = tx_refs + (size_t)tx * 33;
    uint64_t v = 0;
    if (w < WIT_DELTA) v = W[w];
    else if (w == WIT_DELTA) v = W[28];
    else if (w < WIT_S_IDX) v = version(hashes, defaults, R[32])[w - WIT_ROOT];
    else if (w == WIT_S_IDX) v = W[29];
    else if (w == WIT_R_IDX) v = W[30];
    else if (w < WIT_RX) {
        const unsigned k = (w - WIT_S_PATH) / 7, i = (w - WIT_S_PATH) % 7;      // 32 path slots: sender 0..15, receiver 16..31
        v = version(hashes, defaults, R[k])[i];      // slots above the tree's depth refer to the all-zero digest
    } else if (w < WIT_S) v = sigs[(size_t)tx * 14 + (w - WIT_RX)];
    else if (w < WIT_H) v = sigs[(size_t)tx * 14 + 6 + (w - WIT_S)];
    else if (w < WIT_M26) v = sigs[(size_t)tx * 14 + 10 + (w - WIT_H)];
    else v = 0;                                                  // the last two message words of a transfer
    records[g] = v;
}

}  // namespace

void batch_defaults(unsigned depth, fe *defaults_dev, Stream &st) { CSG_LAUNCH(st, batch_defaults_kernel, 1, 1, 0, depth, defaults_dev); }
void batch_gtable(fe *table_dev, Stream &st) { CSG_LAUNCH(st, batch_gtable_kernel, 16, 64, 0, table_dev); }

void batch_build(const BatchDevice &B, Stream &st) {
    const unsigned T = 64;
    CSG_LAUNCH(st, batch_leaf_kernel, (B.level_off[1] + T - 1) / T, T, 0, B.accounts, B.level_off[1], B.hashes);
    for (unsigned l = 1; l <= B.depth; l++) {
        const unsigned first = B.level_off[l], count = B.level_off[l + 1] - first;
        CSG_LAUNCH(st, batch_merge_kernel, (count + T - 1) / T, T, 0, first, count, B.left, B.right, B.hashes, (const fe *)B.defaults);
    }
    CSG_LAUNCH(st, batch_sign_kernel, (B.ntx + 31) / 32, 32, 0, B.tx_words, B.ntx, (const fe *)B.gtable, B.sigs);
    const unsigned words = B.ntx * WIT_WORDS;
    CSG_LAUNCH(st, batch_pack_kernel, (words + 255) / 256, 256, 0, B.tx_words, B.tx_refs, B.ntx, (const fe *)B.hashes, (const fe *)B.defaults, (const uint64_t *)B.sigs, B.records);
}

}  // namespace csg
