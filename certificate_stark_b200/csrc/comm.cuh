// Collectives of the coset-sharded proof: one proof split over G GPUs by LDE coset.
//
// The exchanges of the split: all-gathers of equal slices (trace coefficients by column block, per-coset composition
// interpolants, DEEP evaluations, subtree roots), one all-to-all (leaf digests into contiguous leaf ranges, so that every GPU
// builds the Merkle subtree of its range) and sums of vectors with disjoint support (opened rows and path nodes).
// Two transports implement them:
//   * NcclComm  -- one process per GPU (torchrun): NCCL over NVLink/NVSwitch on the proving stream.  libnccl is
//                  dlopen'ed when the communicator is created, so single-GPU deployments need no NCCL at all.
//   * LocalComm -- one process driving several contexts from one thread each (a Rust host with a thread per GPU, or the
//                  tests, where all "ranks" may even share one device): peer copies + host barriers.
#pragma once
#include <memory>
#include <vector>

#include "dev.cuh"

namespace csg {

struct Comm {
    int rank = 0, world = 1;
    virtual ~Comm() {}
    // in place: buf holds `world` slices of `bytes`; slice `rank` has been produced on stream s; on return (stream order)
    // every slice is filled
    virtual void all_gather(void *buf, size_t bytes, Stream &st) = 0;
    // `bytes` at buf, produced on stream st by rank `root`, arrive at the same address on every other rank (stream order).
    // An all-gather issued slice by slice: the receivers can start on slice q while slice q+1 is still on the links.
    virtual void broadcast(void *buf, size_t bytes, int root, Stream &st) = 0;
    // recv slice p <- slice `rank` of peer p's send buffer; `world` slices of `bytes` each on both sides (send != recv).
    // The digest exchange of a commitment: every rank hashed the rows of its cosets and needs the leaves of a contiguous range.
    virtual void all_to_all(const void *send, void *recv, size_t bytes, Stream &st) = 0;
    // buf[i] = sum over ranks of buf[i]  (u64 wrap-around; callers keep the supports disjoint)
    virtual void all_reduce_sum_u64(uint64_t *buf, size_t count, Stream &st) = 0;
    virtual const char *transport() const = 0;
};

// 128-byte NCCL unique id (rank 0 creates it, the host distributes it to the other ranks)
void nccl_unique_id(uint8_t id[128]);
std::unique_ptr<Comm> make_nccl_comm(int rank, int world, const uint8_t id[128]);
// `world` communicators sharing one in-process group; element r is rank r's
std::vector<std::unique_ptr<Comm>> make_local_comms(int world);

// natural order from per-rank slices: out[(b*i + r*bl + kl)*E + e] = in[((r*n + i)*bl + kl)*E + e], b = G*bl, E u64 words
// per element (4 for a digest).  in: what all_gather leaves when every rank wrote its rows as [i][kl].
void interleave_slices(const uint64_t *in, uint64_t *out, size_t n, unsigned bl, unsigned G, unsigned E, Stream &st);

}  // namespace csg
