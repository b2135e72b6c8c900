// Transports of comm.cuh: NCCL (dlopen'ed) and the in-process thread group.
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "comm.cuh"

namespace csg {
namespace {

// ---------------------------------------------------------------------------------------------- NCCL through dlopen
// The handful of NCCL entry points used here, with the ABI of nccl.h 2.x (ncclUniqueId = 128 bytes by value,
// ncclUint64 = 5, ncclUint8 = 1, ncclSum = 0).
struct NcclId { char internal[128]; };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_SUM = 0;

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("CSG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};   // a process that imported torch already has its libnccl.so.2 mapped
        for (const char *nm : names) {
            if (!nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        auto sym = [&](const char *s) { return dlsym(api.handle, s); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce || !api.Broadcast || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
        throw std::runtime_error("NCCL is not available: libnccl.so.2 could not be loaded (set CSG_NCCL_LIB to its path)");
    return api;
}
void nccl_check(int rc, const char *what) {
    if (rc != 0) throw CudaError(std::string(what) + " failed: " + nccl().GetErrorString(rc));
}

struct NcclComm final : Comm {
    void *comm = nullptr;
    ~NcclComm() override { if (comm) nccl().CommDestroy(comm); }
    void all_gather(void *buf, size_t bytes, Stream &st) override {
        nccl_check(nccl().AllGather((const char *)buf + (size_t)rank * bytes, buf, bytes, NCCL_UINT8, comm, st.s), "ncclAllGather");
    }
    void broadcast(void *buf, size_t bytes, int root, Stream &st) override {
        nccl_check(nccl().Broadcast(buf, buf, bytes, NCCL_UINT8, root, comm, st.s), "ncclBroadcast");
    }
    void all_to_all(const void *send, void *recv, size_t bytes, Stream &st) override {
        nccl_check(nccl().GroupStart(), "ncclGroupStart");
        for (int p = 0; p < world; p++) {
            nccl_check(nccl().Send((const char *)send + (size_t)p * bytes, bytes, NCCL_UINT8, p, comm, st.s), "ncclSend");
            nccl_check(nccl().Recv((char *)recv + (size_t)p * bytes, bytes, NCCL_UINT8, p, comm, st.s), "ncclRecv");
        }
        nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
    }
    void all_reduce_sum_u64(uint64_t *buf, size_t count, Stream &st) override {
        nccl_check(nccl().AllReduce(buf, buf, count, NCCL_UINT64, NCCL_SUM, comm, st.s), "ncclAllReduce");
    }
    const char *transport() const override { return "nccl"; }
};

// ---------------------------------------------------------------------------------------------- in-process group
struct LocalGroup {
    int world = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long generation = 0;
    std::vector<void *> ptr;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long long g = generation;
        if (++arrived == world) { arrived = 0; generation++; cv.notify_all(); }
        else if (!cv.wait_for(lk, std::chrono::seconds(300), [&] { return generation != g; }))
            throw std::runtime_error("local group barrier timed out: a peer rank left the proof (see its error)");
    }
};

__global__ void add_u64_kernel(uint64_t *__restrict__ dst, const uint64_t *__restrict__ src, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < count) dst[i] += src[i];
}

struct LocalComm final : Comm {
    std::shared_ptr<LocalGroup> grp;
    DBuf<uint64_t> tmp;
    void all_gather(void *buf, size_t bytes, Stream &st) override {
        CSG_CUDA(cudaStreamSynchronize(st.s));           // own slice is complete
        grp->ptr[rank] = buf;
        grp->barrier();
        for (int p = 0; p < world; p++)
            if (p != rank)
                CSG_CUDA(cudaMemcpyAsync((char *)buf + (size_t)p * bytes, (const char *)grp->ptr[p] + (size_t)p * bytes, bytes, cudaMemcpyDefault, st.s));
        CSG_CUDA(cudaStreamSynchronize(st.s));
        grp->barrier();                                  // nobody reuses a buffer a peer is still reading
    }
    void broadcast(void *buf, size_t bytes, int root, Stream &st) override {
        CSG_CUDA(cudaStreamSynchronize(st.s));           // the root's slice is complete, the receivers' buffers are free
        grp->ptr[rank] = buf;
        grp->barrier();
        if (rank != root) CSG_CUDA(cudaMemcpyAsync(buf, grp->ptr[root], bytes, cudaMemcpyDefault, st.s));
        CSG_CUDA(cudaStreamSynchronize(st.s));
        grp->barrier();
    }
    void all_to_all(const void *send, void *recv, size_t bytes, Stream &st) override {
        CSG_CUDA(cudaStreamSynchronize(st.s));           // own send buffer is complete
        grp->ptr[rank] = const_cast<void *>(send);
        grp->barrier();
        for (int p = 0; p < world; p++)
            CSG_CUDA(cudaMemcpyAsync((char *)recv + (size_t)p * bytes, (const char *)grp->ptr[p] + (size_t)rank * bytes, bytes, cudaMemcpyDefault, st.s));
        CSG_CUDA(cudaStreamSynchronize(st.s));
        grp->barrier();
    }
    void all_reduce_sum_u64(uint64_t *buf, size_t count, Stream &st) override {
        tmp.reserve(count * (size_t)world);
        CSG_CUDA(cudaStreamSynchronize(st.s));
        grp->ptr[rank] = buf;
        grp->barrier();
        for (int p = 0; p < world; p++)
            if (p != rank) CSG_CUDA(cudaMemcpyAsync(tmp.p + (size_t)p * count, grp->ptr[p], count * 8, cudaMemcpyDefault, st.s));
        CSG_CUDA(cudaStreamSynchronize(st.s));
        grp->barrier();
        for (int p = 0; p < world; p++)
            if (p != rank) CSG_LAUNCH(st, add_u64_kernel, (unsigned)((count + 255) / 256), 256, 0, buf, (const uint64_t *)(tmp.p + (size_t)p * count), count);
    }
    const char *transport() const override { return "local"; }
};

__global__ void interleave_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, unsigned long long n, unsigned bl, unsigned G, unsigned E,
                                  unsigned long long total) {
    const unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const unsigned e = (unsigned)(idx % E);
    const unsigned long long j = idx / E, b = (unsigned long long)bl * G, i = j / b;
    const unsigned k = (unsigned)(j % b), r = k / bl, kl = k % bl;
    out[idx] = in[(((unsigned long long)r * n + i) * bl + kl) * E + e];
}

}  // namespace

void nccl_unique_id(uint8_t id[128]) {
    NcclId u;
    nccl_check(nccl().GetUniqueId(&u), "ncclGetUniqueId");
    memcpy(id, u.internal, 128);
}

std::unique_ptr<Comm> make_nccl_comm(int rank, int world, const uint8_t id[128]) {
    std::unique_ptr<NcclComm> c(new NcclComm());
    c->rank = rank; c->world = world;
    NcclId u;
    memcpy(u.internal, id, 128);
    nccl_check(nccl().CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank");
    return c;
}

std::vector<std::unique_ptr<Comm>> make_local_comms(int world) {
    auto grp = std::make_shared<LocalGroup>();
    grp->world = world;
    grp->ptr.assign(world, nullptr);
    std::vector<std::unique_ptr<Comm>> out;
    for (int r = 0; r < world; r++) {
        std::unique_ptr<LocalComm> c(new LocalComm());
        c->rank = r; c->world = world; c->grp = grp;
        out.push_back(std::move(c));
    }
    return out;
}

void interleave_slices(const uint64_t *in, uint64_t *out, size_t n, unsigned bl, unsigned G, unsigned E, Stream &st) {
    const unsigned long long total = (unsigned long long)n * bl * G * E;
    CSG_LAUNCH(st, interleave_kernel, (unsigned)((total + 255) / 256), 256, 0, in, out, (unsigned long long)n, bl, G, E, total);
}

}  // namespace csg
