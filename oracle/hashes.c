/* ORACLE -- TEST INFRASTRUCTURE ONLY.  BLAKE3 (default 256-bit output, unkeyed) and SHA3-256 from their specs. */
#include "hashes.h"
#include <string.h>

/* ------------------------------------------------------------------ BLAKE3 */
static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

static inline uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
#define B3_G(a, b, c, d, x, y)                                             \
    do {                                                                   \
        v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);          \
        v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 12);          \
        v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);           \
        v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 7);           \
    } while (0)

/* out[8] = first half of the compression output (the chaining value / first 32 output bytes) */
static void b3_compress(const uint32_t cv[8], const uint8_t block[64], uint32_t block_len, uint64_t counter,
                        uint32_t flags, uint32_t out[8]) {
    uint32_t m[16], v[16], t[16];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) | ((uint32_t)block[4 * i + 2] << 16) |
               ((uint32_t)block[4 * i + 3] << 24);
    for (int i = 0; i < 8; i++) v[i] = cv[i];
    v[8] = B3_IV[0]; v[9] = B3_IV[1]; v[10] = B3_IV[2]; v[11] = B3_IV[3];
    v[12] = (uint32_t)counter; v[13] = (uint32_t)(counter >> 32); v[14] = block_len; v[15] = flags;
    for (int r = 0; r < 7; r++) {
        B3_G(0, 4, 8, 12, m[0], m[1]);   B3_G(1, 5, 9, 13, m[2], m[3]);
        B3_G(2, 6, 10, 14, m[4], m[5]);  B3_G(3, 7, 11, 15, m[6], m[7]);
        B3_G(0, 5, 10, 15, m[8], m[9]);  B3_G(1, 6, 11, 12, m[10], m[11]);
        B3_G(2, 7, 8, 13, m[12], m[13]); B3_G(3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[B3_PERM[i]];
        memcpy(m, t, sizeof m);
    }
    for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}

/* chaining value of one chunk (<= 1024 bytes); `root` adds the ROOT flag on its last block */
static void b3_chunk(const uint8_t *in, size_t len, uint64_t chunk_counter, int root, uint32_t out[8]) {
    uint32_t cv[8];
    memcpy(cv, B3_IV, sizeof cv);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t block[64] = {0};
        size_t off = b * 64, bl = len - off < 64 ? len - off : 64;
        memcpy(block, in + off, bl);
        uint32_t flags = 0;
        if (b == 0) flags |= B3_CHUNK_START;
        if (b == nblocks - 1) flags |= B3_CHUNK_END | (root ? B3_ROOT : 0);
        b3_compress(cv, block, (uint32_t)bl, chunk_counter, flags, cv);
    }
    memcpy(out, cv, sizeof cv);
}
/* chaining value of a subtree covering `len` bytes (> 0) that starts at chunk index `chunk_counter` */
static void b3_subtree(const uint8_t *in, size_t len, uint64_t chunk_counter, int root, uint32_t out[8]) {
    if (len <= 1024) { b3_chunk(in, len, chunk_counter, root, out); return; }
    /* left subtree: the largest power-of-two number of chunks that leaves at least one byte on the right */
    size_t nchunks = (len + 1023) / 1024, left = 1;
    while (left * 2 < nchunks) left *= 2;
    uint32_t l[8], r[8];
    b3_subtree(in, left * 1024, chunk_counter, 0, l);
    b3_subtree(in + left * 1024, len - left * 1024, chunk_counter + left, 0, r);
    uint8_t block[64];
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 4; k++) { block[4 * i + k] = (uint8_t)(l[i] >> (8 * k)); block[32 + 4 * i + k] = (uint8_t)(r[i] >> (8 * k)); }
    b3_compress(B3_IV, block, 64, 0, B3_PARENT | (root ? B3_ROOT : 0), out);
}
void blake3_256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t cv[8];
    b3_subtree(in, len, 0, 1, cv);
    for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(cv[i] >> (8 * k));
}

/* ------------------------------------------------------------------ SHA3-256 (Keccak-f[1600], rate 136, pad 0x06..0x80) */
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t a[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        /* rho + pi: B[y][2x+3y] = rot(A[x][y]) with index = x + 5y */
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(a[x + 5 * y], KECCAK_ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= KECCAK_RC[round];
    }
}
void sha3_256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint64_t st[25] = {0};
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, in + 8 * i, 8); st[i] ^= w; }
        keccak_f(st);
        in += rate; len -= rate;
    }
    uint8_t last[136] = {0};
    memcpy(last, in, len);
    last[len] ^= 0x06; last[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, last + 8 * i, 8); st[i] ^= w; }
    keccak_f(st);
    memcpy(out, st, 32);
}

void hash_bytes(int hash_fn, const uint8_t *in, size_t len, uint8_t out[32]) {
    if (hash_fn == HASH_SHA3_256) sha3_256(in, len, out); else blake3_256(in, len, out);
}
