// Plan of a transaction batch for the device-side builder (batch_gen.cu): everything TransactionMetadata::build_random
// (/root/reference/src/lib.rs:235-464) decides WITHOUT hashing -- the seeded draws (accounts, indices, amounts), the account values
// before and after every transfer, and the shape of the Rescue Merkle tree's history: which node versions exist and which two child
// versions each one merges.  The hashes themselves (leaves, every node version, signature hashes) are computed by kernels.
//
// The tree after the initial accounts is "time 0"; transfer t updates the sender's leaf at time 2t+1 and the receiver's at 2t+2.
// A record is one version of one node.  Level l (0 = leaves) has nbase[l] records for time 0 followed by one record per update
// event; record ids are global (level after level).  A child reference is a record id, or -(level+1) for the empty subtree of that
// level (winterfell MerkleTree::build_empty: all-zero leaves).
#pragma once
#include <cstdint>
#include <vector>

namespace csg {

struct BatchPlan {
    unsigned depth = 0;
    size_t ntx = 0;
    std::vector<uint32_t> level_off;    // first record id of each level 0..depth, then the total
    std::vector<uint32_t> nbase;        // time-0 records per level
    std::vector<uint64_t> accounts;     // level-0 records: 14 Montgomery words each (key[12], balance, nonce)
    std::vector<int32_t> left, right;   // per record id >= level_off[1]: the child versions it merges
    enum : int { TX_WORDS = 33, TX_REFS = 33 };
    std::vector<uint64_t> tx_words;     // per transfer: s_old[14] r_old[14] delta s_idx r_idx sk sig_seed
    std::vector<int32_t> tx_refs;       // per transfer: sender path [16] (leaf, then the sibling at each level), receiver path [16], root before the transfer
    int32_t final_root = 0;
};
// throws std::invalid_argument on bad shapes; the draws are exactly those of csg_tx_batch_new(seed, num_tx, tree_depth)
BatchPlan plan_tx_batch(uint64_t seed, size_t num_tx, unsigned tree_depth);

}  // namespace csg
