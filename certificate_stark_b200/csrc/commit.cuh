// Vector commitments (K3/K4 of SURVEY.md section 2.2): leaf = H(row bytes), node = H(left || right), with the hash the
// ProofOptions select -- what winterfell's `H::hash_elements(row)` + `MerkleTree::new(leaves)` do for the reference
// (stage 2 and 4 of Prover::prove, entered at /root/reference/src/lib.rs:140).
//
// Layout: a tree over L leaves is one array of 2L digests of 8 u32 words; nodes[1] is the root, nodes[i] has children
// nodes[2i], nodes[2i+1], leaf j sits at nodes[L + j] -- the row-hash kernel writes leaves straight into that slot.
#pragma once
#include "dev.cuh"

namespace csg {

// digest of row j = k + ncosets*i of a coset-major matrix: elements data[k*coset_stride + c*col_stride + i], c < width,
// hashed as canonical little-endian bytes.  Writes 8 words to leaves + 8*j.
// sub > 1: column c lives at (c / sub) * col_stride + (c % sub) * sub_stride (extension-field elements stored as planes).
void hash_rows(const fe *data, unsigned width, size_t n, unsigned ncosets, size_t coset_stride, size_t col_stride, int hash_fn,
               uint32_t *leaves, Stream &st, unsigned sub = 1, size_t sub_stride = 0);
// interior nodes of the tree whose leaves are already in nodes[8*L ..)
void merkle_build(uint32_t *nodes, size_t nleaves, int hash_fn, Stream &st);
// out[8*t ..] = nodes[8*idx[t] ..]; for the per-GPU subtrees of a sharded proof idx may also name a node of `top`
// (0x80000000 | i) or a node another rank owns (0xFFFFFFFF: zeros)
void gather_digests(const uint32_t *nodes, const uint32_t *idx_dev, size_t count, uint32_t *out_dev, Stream &st, const uint32_t *top = nullptr);

}  // namespace csg
