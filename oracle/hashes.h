/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * The two proof hashes the reference can select (src/lib.rs:82 HashFunction::Blake3_256,
 * examples/state-transition.rs:67-71 Sha3_256).  The reference calls them through winterfell's
 * Blake3_256 / Sha3_256 hashers, which wrap the `blake3` and `sha3` crates (absent here); both are
 * restated from their published specifications and pinned in tests against Python's `blake3` module
 * and hashlib.sha3_256.
 */
#ifndef ORACLE_HASHES_H
#define ORACLE_HASHES_H
#include <stdint.h>
#include <stddef.h>

enum { HASH_BLAKE3_192 = 1, HASH_BLAKE3_256 = 2, HASH_SHA3_256 = 3 }; /* winterfell HashFunction repr [RECALLED] */

void blake3_256(const uint8_t *in, size_t len, uint8_t out[32]);
void sha3_256(const uint8_t *in, size_t len, uint8_t out[32]);
/* dispatch on the proof option */
void hash_bytes(int hash_fn, const uint8_t *in, size_t len, uint8_t out[32]);
#endif
