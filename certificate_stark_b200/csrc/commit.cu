// See commit.cuh.  One thread = one message; message words come straight from coalesced column reads.
#include "commit.cuh"
#include "hash.cuh"

namespace csg {
using namespace f63;

namespace {

// SUB: column c lives at (c / sub) * col_stride + (c % sub) * sub_stride (rows of extension-field elements stored as planes)
template <int HASH, bool SUB>
__global__ void __launch_bounds__(128) hash_rows_kernel(const fe *__restrict__ data, unsigned width, unsigned long long n, unsigned ncosets,
                                                        unsigned long long coset_stride, unsigned long long col_stride,
                                                        uint32_t *__restrict__ leaves, unsigned sub, unsigned long long sub_stride) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = blockIdx.y;
    const fe *p = data + k * coset_stride + i;
    uint32_t d[8];
    auto get = [&](uint32_t c) -> uint64_t {
        if (SUB) return from_mont(__ldg(p + (c / sub) * col_stride + (c % sub) * sub_stride));
        return from_mont(__ldg(p + c * col_stride));
    };
    if (HASH == hashes::SHA3_256) hashes::k3::hash_words64(get, width, d); else hashes::b3::hash_words64(get, width, d);
    uint4 *o = reinterpret_cast<uint4 *>(leaves + 8ULL * (k + (unsigned long long)ncosets * i));
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

__device__ __forceinline__ void load_digest(const uint32_t *p, uint32_t (&d)[8]) {
    uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
__device__ __forceinline__ void store_digest(uint32_t *p, const uint32_t (&d)[8]) {
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// one level: nodes[i] = H(nodes[2i] || nodes[2i+1]) for i in [m, 2m)
template <int HASH>
__global__ void __launch_bounds__(256) merkle_level_kernel(uint32_t *nodes, unsigned long long m) {
    unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (t >= m) return;
    unsigned long long i = m + t;
    uint32_t l[8], r[8], o[8];
    load_digest(nodes + 16 * i, l);
    load_digest(nodes + 16 * i + 8, r);
    hashes::merge(HASH, l, r, o);
    store_digest(nodes + 8 * i, o);
}
// the top of the tree (levels of at most 512 nodes) in one CTA
template <int HASH>
__global__ void __launch_bounds__(512) merkle_top_kernel(uint32_t *nodes, unsigned m_first) {
    for (unsigned m = m_first; m >= 1; m >>= 1) {
        if (threadIdx.x < m) {
            unsigned i = m + threadIdx.x;
            uint32_t l[8], r[8], o[8];
            load_digest(nodes + 16 * i, l);
            load_digest(nodes + 16 * i + 8, r);
            hashes::merge(HASH, l, r, o);
            store_digest(nodes + 8 * i, o);
        }
        __syncthreads();
    }
}
// idx: a node of `nodes`; 0x80000000 | i: node i of `top` (the replicated top levels of a sharded tree); 0xFFFFFFFF: a node another
// rank owns -- zeros, the digests of all ranks are summed afterwards
__global__ void gather_digests_kernel(const uint32_t *nodes, const uint32_t *top, const uint32_t *idx, unsigned count, uint32_t *out) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * 8) return;
    const uint32_t i = idx[t >> 3];
    out[t] = i == 0xFFFFFFFFu ? 0u : (i & 0x80000000u) ? top[8ULL * (i & 0x7FFFFFFFu) + (t & 7)] : nodes[8ULL * i + (t & 7)];
}

}  // namespace

void hash_rows(const fe *data, unsigned width, size_t n, unsigned ncosets, size_t coset_stride, size_t col_stride, int hash_fn,
               uint32_t *leaves, Stream &st, unsigned sub, size_t sub_stride) {
    if (width > 128) throw std::runtime_error("rows wider than one Blake3 chunk are not supported");
    dim3 grid((unsigned)((n + 127) / 128), ncosets);
#define CSG_HASH_ROWS(H, S) CSG_LAUNCH(st, (hash_rows_kernel<H, S>), grid, 128, 0, data, width, (unsigned long long)n, ncosets, \
                                       (unsigned long long)coset_stride, (unsigned long long)col_stride, leaves, sub, (unsigned long long)sub_stride)
    if (hash_fn == hashes::SHA3_256) { if (sub > 1) CSG_HASH_ROWS(hashes::SHA3_256, true); else CSG_HASH_ROWS(hashes::SHA3_256, false); }
    else { if (sub > 1) CSG_HASH_ROWS(hashes::BLAKE3_256, true); else CSG_HASH_ROWS(hashes::BLAKE3_256, false); }
#undef CSG_HASH_ROWS
}

void merkle_build(uint32_t *nodes, size_t nleaves, int hash_fn, Stream &st) {
    if (nleaves < 2) throw std::runtime_error("a Merkle tree needs at least two leaves");
    CSG_CUDA(cudaMemsetAsync(nodes, 0, 32, st.s));  // nodes[0] is unused
    size_t m = nleaves / 2;
    for (; m > 512; m >>= 1) {
        unsigned grid = (unsigned)((m + 255) / 256);
        if (hash_fn == hashes::SHA3_256) CSG_LAUNCH(st, merkle_level_kernel<hashes::SHA3_256>, grid, 256, 0, nodes, (unsigned long long)m);
        else CSG_LAUNCH(st, merkle_level_kernel<hashes::BLAKE3_256>, grid, 256, 0, nodes, (unsigned long long)m);
    }
    if (hash_fn == hashes::SHA3_256) CSG_LAUNCH(st, merkle_top_kernel<hashes::SHA3_256>, 1, 512, 0, nodes, (unsigned)m);
    else CSG_LAUNCH(st, merkle_top_kernel<hashes::BLAKE3_256>, 1, 512, 0, nodes, (unsigned)m);
}

void gather_digests(const uint32_t *nodes, const uint32_t *idx_dev, size_t count, uint32_t *out_dev, Stream &st, const uint32_t *top) {
    if (!count) return;
    CSG_LAUNCH(st, gather_digests_kernel, (unsigned)((count * 8 + 255) / 256), 256, 0, nodes, top, idx_dev, (unsigned)count, out_dev);
}

}  // namespace csg
