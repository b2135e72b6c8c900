// Device-side witness generation (witness_gen.cu) and the packed per-transaction input record both sides agree on.
#pragma once
#include "dev.cuh"

namespace csg {

// one record of WIT_WORDS u64 per transaction, field elements in Montgomery form (host: csg_tx_batch_pack, witness.cpp)
enum : int {
    WIT_S_OLD = 0,      // sender leaf before the transfer: key[12], balance, nonce
    WIT_R_OLD = 14,     // receiver leaf
    WIT_DELTA = 28,
    WIT_ROOT = 29,      // tree root before this transaction [7]
    WIT_S_IDX = 36, WIT_R_IDX = 37,
    WIT_S_PATH = 38,    // authentication path of the sender: 16 nodes of 7 elements, node k+1 = sibling at level k
    WIT_R_PATH = 150,
    WIT_RX = 262,       // x coordinate of the signature's R [6]
    WIT_S = 268,        // signature scalar s, 4 little-endian words
    WIT_H = 272,        // message hash h as 4 little-endian words (the bits driving h.P)
    WIT_M26 = 276,      // last two words of the signed message: zero for a transfer (src/lib.rs:467-481), random for SchnorrExample's
    WIT_WORDS = 278     //   messages (src/schnorr/mod.rs:97-99).  The message is S_OLD[0..12] | R_OLD[0..12] | DELTA | S_OLD[13] | M26 | M27
};

// canonical column-major trace of ntx transactions (94 x 1024*ntx) into trace_dev; finals_dev: 48 elements per transaction
void build_transaction_trace(const uint64_t *inputs_dev, size_t ntx, unsigned tree_depth, uint64_t *trace_dev, fe *finals_dev, Stream &st);
// the standalone provers of the sub-AIRs from the same records and the same kernels:
//   MerkleProver::build_trace  (src/merkle/update/prover.rs:37-80): 65 x 512*ntx, rows 0..511 of every transfer's Merkle phase,
//                              plus the two bit cells the reference sets at step 1 (:72-77)
//   SchnorrProver::build_trace (src/schnorr/prover.rs:52-80): 56 x 512*nsig, the signature phase alone
void build_merkle_update_trace(const uint64_t *inputs_dev, size_t ntx, unsigned tree_depth, uint64_t *trace_dev, Stream &st);
void build_schnorr_trace(const uint64_t *inputs_dev, size_t nsig, uint64_t *trace_dev, fe *finals_dev, Stream &st);

// ---- TransactionMetadata::build_random on the device (batch_gen.cu; the plan comes from host/batch_plan.hpp)
struct BatchDevice {
    unsigned depth, ntx;
    const unsigned *level_off;   // HOST array: first record id of each level 0..depth, then the total
    const uint64_t *accounts;    // device: 14 words per level-0 record
    const int *left, *right;     // device: child versions of every record
    const uint64_t *tx_words;    // device: 33 words per transfer
    const int *tx_refs;          // device: 33 version references per transfer
    fe *hashes;                  // device: 7 elements per record (output)
    const fe *defaults;          // device: empty-subtree hash of every level
    const fe *gtable;            // device: fixed-base table of the generator
    uint64_t *sigs;              // device: 14 words per transfer (rx, s, h)
    uint64_t *records;           // device: WIT_WORDS words per transfer (output: what build_transaction_trace reads)
};
void batch_defaults(unsigned depth, fe *defaults_dev, Stream &st);   // (depth + 1) * 7 elements
void batch_gtable(fe *table_dev, Stream &st);                        // 64 * 16 * 12 elements
void batch_build(const BatchDevice &B, Stream &st);

}  // namespace csg
