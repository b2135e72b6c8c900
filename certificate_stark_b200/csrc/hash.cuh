// Blake3-256 and SHA3-256, the two proof hashes the reference can select through ProofOptions
// (HashFunction::Blake3_256 at src/lib.rs:82; Sha3_256 via `-h 1`, examples/state-transition.rs:67-71).
// winterfell's hashers wrap the `blake3` and `sha3` crates; both algorithms are written here from their public
// specifications, as host+device code: the device kernels hash LDE rows / Merkle nodes / FRI rows, the host side runs
// the Fiat-Shamir transcript (RandomCoin) with the very same functions.
//
// Device usage is "one thread = one message": messages on this path are short (a 752-byte row at most), so the state
// lives in registers and the message words are produced on the fly from field elements (canonical little-endian).
#pragma once
#include "field.cuh"

namespace hashes {

enum : int { BLAKE3_256 = 2, SHA3_256 = 3 };  // winterfell HashFunction discriminants used by ProofOptions serialisation

// ------------------------------------------------------------------------------------------------ Blake3
namespace b3 {
constexpr uint32_t CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8;
constexpr uint32_t BLOCK_LEN = 64, CHUNK_LEN = 1024;

#define CSG_B3_IV0 0x6A09E667u
#define CSG_B3_IV1 0xBB67AE85u
#define CSG_B3_IV2 0x3C6EF372u
#define CSG_B3_IV3 0xA54FF53Au
#define CSG_B3_IV4 0x510E527Fu
#define CSG_B3_IV5 0x9B05688Cu
#define CSG_B3_IV6 0x1F83D9ABu
#define CSG_B3_IV7 0x5BE0CD19u

CSG_HD uint32_t rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

// The quarter-round is 4 additions, 4 XORs and 4 rotations: on the GPU all twelve would issue on the ALU pipe (IADD3, LOP3,
// SHF/PRMT), which the row-hash and Merkle kernels saturate (ncu: ALU 85-92 %, FMA 10-13 %).  Writing x + y as x * ONE + y
// with ONE read from constant memory (not foldable at compile time) turns the additions into IMADs on the otherwise idle FMA
// pipe: 8 ALU + 6 FMA instructions per quarter-round instead of 12 ALU.
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t CSG_B3_ONE = 1;
#endif
#if defined(__CUDA_ARCH__)
#define CSG_B3_ADD(x, y) ((x) * CSG_B3_ONE + (y))
#else
#define CSG_B3_ADD(x, y) ((x) + (y))
#endif
#define CSG_B3_G(a, b, c, d, mx, my) \
    a = CSG_B3_ADD(CSG_B3_ADD(a, b), (mx)); d = rotr(d ^ a, 16); c = CSG_B3_ADD(c, d); b = rotr(b ^ c, 12); \
    a = CSG_B3_ADD(CSG_B3_ADD(a, b), (my)); d = rotr(d ^ a, 8);  c = CSG_B3_ADD(c, d); b = rotr(b ^ c, 7);

#define CSG_B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    CSG_B3_G(s0, s4, s8, s12, m0, m1)  CSG_B3_G(s1, s5, s9, s13, m2, m3)                     \
    CSG_B3_G(s2, s6, s10, s14, m4, m5) CSG_B3_G(s3, s7, s11, s15, m6, m7)                    \
    CSG_B3_G(s0, s5, s10, s15, m8, m9) CSG_B3_G(s1, s6, s11, s12, m10, m11)                  \
    CSG_B3_G(s2, s7, s8, s13, m12, m13) CSG_B3_G(s3, s4, s9, s14, m14, m15)

// One compression.  cv is updated in place to the new chaining value (first 8 output words).  The message schedule
// (the fixed permutation applied between rounds) is expanded at compile time, so m[] stays in registers.
CSG_HD void compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint64_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = CSG_B3_IV0, s9 = CSG_B3_IV1, s10 = CSG_B3_IV2, s11 = CSG_B3_IV3;
    uint32_t s12 = (uint32_t)counter, s13 = (uint32_t)(counter >> 32), s14 = block_len, s15 = flags;
    CSG_B3_ROUND(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15])
    CSG_B3_ROUND(m[2], m[6], m[3], m[10], m[7], m[0], m[4], m[13], m[1], m[11], m[12], m[5], m[9], m[14], m[15], m[8])
    CSG_B3_ROUND(m[3], m[4], m[10], m[12], m[13], m[2], m[7], m[14], m[6], m[5], m[9], m[0], m[11], m[15], m[8], m[1])
    CSG_B3_ROUND(m[10], m[7], m[12], m[9], m[14], m[3], m[13], m[15], m[4], m[0], m[11], m[2], m[5], m[8], m[1], m[6])
    CSG_B3_ROUND(m[12], m[13], m[9], m[11], m[15], m[10], m[14], m[8], m[7], m[2], m[5], m[3], m[0], m[1], m[6], m[4])
    CSG_B3_ROUND(m[9], m[14], m[11], m[5], m[8], m[12], m[15], m[1], m[13], m[3], m[0], m[10], m[2], m[6], m[4], m[7])
    CSG_B3_ROUND(m[11], m[15], m[5], m[0], m[1], m[9], m[8], m[6], m[14], m[10], m[2], m[12], m[3], m[4], m[7], m[13])
    cv[0] = s0 ^ s8; cv[1] = s1 ^ s9; cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}
CSG_HD void iv(uint32_t (&cv)[8]) {
    cv[0] = CSG_B3_IV0; cv[1] = CSG_B3_IV1; cv[2] = CSG_B3_IV2; cv[3] = CSG_B3_IV3;
    cv[4] = CSG_B3_IV4; cv[5] = CSG_B3_IV5; cv[6] = CSG_B3_IV6; cv[7] = CSG_B3_IV7;
}

// Hash of `nwords64` 64-bit little-endian words produced by get(i), for messages of at most one chunk (<= 1024 bytes,
// i.e. nwords64 <= 128): every message the device hashes (LDE rows, composition rows, FRI rows, Merkle node pairs).
template <class Get>
CSG_HD void hash_words64(Get get, uint32_t nwords64, uint32_t (&out)[8]) {
    iv(out);
    uint32_t nblocks = nwords64 == 0 ? 1 : (nwords64 + 7) / 8;
    for (uint32_t b = 0; b < nblocks; b++) {
        uint32_t m[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; j++) {
            uint32_t idx = b * 8 + j;
            uint64_t v = idx < nwords64 ? get(idx) : 0;
            m[2 * j] = (uint32_t)v; m[2 * j + 1] = (uint32_t)(v >> 32);
        }
        uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b + 1 == nblocks ? (CHUNK_END | ROOT) : 0);
        uint32_t len = b + 1 == nblocks ? (nwords64 - b * 8) * 8 : BLOCK_LEN;
        compress(out, m, 0, len, flags);
    }
}
// parent of two 32-byte digests as winterfell's merge() forms it: hash of the 64 concatenated bytes (a one-block chunk)
CSG_HD void merge(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
    uint32_t m[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 8; j++) { m[j] = l[j]; m[8 + j] = r[j]; }
    iv(out);
    compress(out, m, 0, BLOCK_LEN, CHUNK_START | CHUNK_END | ROOT);
}

// ---- host: full Blake3 (arbitrary length, tree mode) for the transcript seed (public inputs can exceed one chunk)
inline void load_block(const uint8_t *p, size_t len, uint32_t (&m)[16]) {
    uint8_t tmp[64] = {0};
    for (size_t i = 0; i < len; i++) tmp[i] = p[i];
    for (int j = 0; j < 16; j++) m[j] = (uint32_t)tmp[4 * j] | ((uint32_t)tmp[4 * j + 1] << 8) | ((uint32_t)tmp[4 * j + 2] << 16) | ((uint32_t)tmp[4 * j + 3] << 24);
}
// chaining value of one chunk; `root` marks the single-chunk message
inline void chunk_cv(const uint8_t *p, size_t len, uint64_t chunk_index, bool root, uint32_t (&cv)[8]) {
    iv(cv);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint32_t m[16];
        size_t bl = b + 1 == nblocks ? len - b * 64 : 64;
        load_block(p + b * 64, bl, m);
        uint32_t flags = (b == 0 ? CHUNK_START : 0) | (b + 1 == nblocks ? (CHUNK_END | (root ? ROOT : 0)) : 0);
        compress(cv, m, chunk_index, (uint32_t)bl, flags);
    }
}
// chaining value of the subtree covering p[0..len), len > CHUNK_LEN possible; left subtree takes the largest power of
// two number of chunks that leaves at least one byte on the right
inline void subtree_cv(const uint8_t *p, size_t len, uint64_t chunk0, bool root, uint32_t (&cv)[8]) {
    if (len <= CHUNK_LEN) { chunk_cv(p, len, chunk0, root, cv); return; }
    size_t chunks = (len - 1) / CHUNK_LEN, left = 1;
    while (left * 2 <= chunks) left *= 2;
    uint32_t l[8], r[8], m[16];
    subtree_cv(p, left * CHUNK_LEN, chunk0, false, l);
    subtree_cv(p + left * CHUNK_LEN, len - left * CHUNK_LEN, chunk0 + left, false, r);
    for (int j = 0; j < 8; j++) { m[j] = l[j]; m[8 + j] = r[j]; }
    iv(cv);
    compress(cv, m, 0, BLOCK_LEN, PARENT | (root ? ROOT : 0));
}
inline void hash_bytes(const uint8_t *p, size_t len, uint8_t out[32]) {
    uint32_t cv[8];
    subtree_cv(p, len, 0, true, cv);
    for (int j = 0; j < 8; j++) for (int k = 0; k < 4; k++) out[4 * j + k] = (uint8_t)(cv[j] >> (8 * k));
}
}  // namespace b3

// ------------------------------------------------------------------------------------------------ SHA3-256 (Keccak-f[1600], rate 136)
namespace k3 {
CSG_HD uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

#define CSG_K3_RC(i)                                                                                                          \
    ((i) == 0 ? 0x0000000000000001ULL : (i) == 1 ? 0x0000000000008082ULL : (i) == 2 ? 0x800000000000808aULL :                 \
     (i) == 3 ? 0x8000000080008000ULL : (i) == 4 ? 0x000000000000808bULL : (i) == 5 ? 0x0000000080000001ULL :                 \
     (i) == 6 ? 0x8000000080008081ULL : (i) == 7 ? 0x8000000000008009ULL : (i) == 8 ? 0x000000000000008aULL :                 \
     (i) == 9 ? 0x0000000000000088ULL : (i) == 10 ? 0x0000000080008009ULL : (i) == 11 ? 0x000000008000000aULL :               \
     (i) == 12 ? 0x000000008000808bULL : (i) == 13 ? 0x800000000000008bULL : (i) == 14 ? 0x8000000000008089ULL :              \
     (i) == 15 ? 0x8000000000008003ULL : (i) == 16 ? 0x8000000000008002ULL : (i) == 17 ? 0x8000000000000080ULL :              \
     (i) == 18 ? 0x000000000000800aULL : (i) == 19 ? 0x800000008000000aULL : (i) == 20 ? 0x8000000080008081ULL :              \
     (i) == 21 ? 0x8000000000008080ULL : (i) == 22 ? 0x0000000080000001ULL : 0x8000000080008008ULL)

CSG_HD void keccak_f(uint64_t (&a)[25]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int round = 0; round < 24; round++) {
        uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20], c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
        uint64_t c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22], c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
        uint64_t c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        uint64_t d0 = c4 ^ rotl(c1, 1), d1 = c0 ^ rotl(c2, 1), d2 = c1 ^ rotl(c3, 1), d3 = c2 ^ rotl(c4, 1), d4 = c3 ^ rotl(c0, 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int y = 0; y < 25; y += 5) { a[y] ^= d0; a[y + 1] ^= d1; a[y + 2] ^= d2; a[y + 3] ^= d3; a[y + 4] ^= d4; }
        // rho + pi
        uint64_t b[25];
        b[0] = a[0];             b[10] = rotl(a[1], 1);   b[20] = rotl(a[2], 62);  b[5] = rotl(a[3], 28);   b[15] = rotl(a[4], 27);
        b[16] = rotl(a[5], 36);  b[1] = rotl(a[6], 44);   b[11] = rotl(a[7], 6);   b[21] = rotl(a[8], 55);  b[6] = rotl(a[9], 20);
        b[7] = rotl(a[10], 3);   b[17] = rotl(a[11], 10); b[2] = rotl(a[12], 43);  b[12] = rotl(a[13], 25); b[22] = rotl(a[14], 39);
        b[23] = rotl(a[15], 41); b[8] = rotl(a[16], 45);  b[18] = rotl(a[17], 15); b[3] = rotl(a[18], 21);  b[13] = rotl(a[19], 8);
        b[14] = rotl(a[20], 18); b[24] = rotl(a[21], 2);  b[9] = rotl(a[22], 61);  b[19] = rotl(a[23], 56); b[4] = rotl(a[24], 14);
        // chi
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int y = 0; y < 25; y += 5) {
            a[y] = b[y] ^ (~b[y + 1] & b[y + 2]);         a[y + 1] = b[y + 1] ^ (~b[y + 2] & b[y + 3]);
            a[y + 2] = b[y + 2] ^ (~b[y + 3] & b[y + 4]); a[y + 3] = b[y + 3] ^ (~b[y + 4] & b[y]);
            a[y + 4] = b[y + 4] ^ (~b[y] & b[y + 1]);
        }
        a[0] ^= CSG_K3_RC(round);
    }
}
// SHA3-256 of nwords64 little-endian 64-bit words produced by get(i) (message length a multiple of 8 bytes)
template <class Get>
CSG_HD void hash_words64(Get get, uint32_t nwords64, uint32_t (&out)[8]) {
    uint64_t a[25];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 25; i++) a[i] = 0;
    const uint32_t RATE_WORDS = 17;  // 136 bytes
    uint32_t pos = 0;
    while (nwords64 - pos >= RATE_WORDS) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 17; i++) a[i] ^= get(pos + i);
        keccak_f(a);
        pos += RATE_WORDS;
    }
    uint32_t rem = nwords64 - pos;  // < 17 whole words remain; padding 0x06 .. 0x80 goes after them
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 17; i++) {
        uint64_t v = (uint32_t)i < rem ? get(pos + i) : 0;
        if ((uint32_t)i == rem) v ^= 0x06ULL;
        if (i == 16) v ^= 0x8000000000000000ULL;
        a[i] ^= v;
    }
    keccak_f(a);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) { out[2 * i] = (uint32_t)a[i]; out[2 * i + 1] = (uint32_t)(a[i] >> 32); }
}
CSG_HD void merge(const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
    uint64_t w[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; j++) { w[j] = (uint64_t)l[2 * j] | ((uint64_t)l[2 * j + 1] << 32); w[4 + j] = (uint64_t)r[2 * j] | ((uint64_t)r[2 * j + 1] << 32); }
    hash_words64([&](uint32_t i) { return w[i]; }, 8, out);
}
inline void hash_bytes(const uint8_t *p, size_t len, uint8_t out[32]) {
    uint64_t a[25] = {0};
    uint8_t blk[136];
    size_t pos = 0;
    auto absorb = [&](const uint8_t *b) {
        for (int i = 0; i < 17; i++) { uint64_t v = 0; for (int k = 0; k < 8; k++) v |= (uint64_t)b[8 * i + k] << (8 * k); a[i] ^= v; }
        keccak_f(a);
    };
    while (len - pos >= 136) { absorb(p + pos); pos += 136; }
    size_t rem = len - pos;
    for (size_t i = 0; i < 136; i++) blk[i] = i < rem ? p[pos + i] : 0;
    blk[rem] ^= 0x06; blk[135] ^= 0x80;
    absorb(blk);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(a[i] >> (8 * k));
}
}  // namespace k3

// ---- dispatch on the ProofOptions hash function
template <class Get>
CSG_HD void hash_words64(int hash_fn, Get get, uint32_t nwords64, uint32_t (&out)[8]) {
    if (hash_fn == SHA3_256) k3::hash_words64(get, nwords64, out); else b3::hash_words64(get, nwords64, out);
}
CSG_HD void merge(int hash_fn, const uint32_t (&l)[8], const uint32_t (&r)[8], uint32_t (&out)[8]) {
    if (hash_fn == SHA3_256) k3::merge(l, r, out); else b3::merge(l, r, out);
}
inline void hash_bytes(int hash_fn, const uint8_t *p, size_t len, uint8_t out[32]) {
    if (hash_fn == SHA3_256) k3::hash_bytes(p, len, out); else b3::hash_bytes(p, len, out);
}
}  // namespace hashes
