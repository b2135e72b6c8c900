// Device plumbing shared by the .cu files of libcsg: error handling, a launch-counting stream wrapper and a
// trivially simple device buffer.  No torch types anywhere: the library is plain CUDA runtime behind a C ABI.
#pragma once
#include <chrono>
#include <utility>
#include <vector>
#include <cuda_runtime.h>

#include <cstdio>
#include <stdexcept>
#include <string>

#include "field.cuh"

namespace csg {
using f63::fe;

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
        throw CudaError(buf);
    }
}
#define CSG_CUDA(x) ::csg::cuda_check((x), #x, __FILE__, __LINE__)

// CSG_HOST_TRACE: host clock marks of the proof in flight on this thread (prover_ctx.cuh prints them); a no-op otherwise
struct HostTrace {
    std::chrono::steady_clock::time_point t0;
    std::vector<std::pair<const char *, double>> marks;
};
inline thread_local HostTrace *g_host_trace = nullptr;
inline void host_mark(const char *what) {
    if (g_host_trace) g_host_trace->marks.emplace_back(what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - g_host_trace->t0).count());
}

// the proving stream; every kernel launch of the library goes through LAUNCH so that launches are counted
struct Stream {
    cudaStream_t s = nullptr;
    unsigned long long launches = 0;
};
#define CSG_LAUNCH(st, kernel, grid, block, smem, ...)                         \
    do {                                                                       \
        kernel<<<(grid), (block), (smem), (st).s>>>(__VA_ARGS__);              \
        (st).launches++;                                                       \
        CSG_CUDA(cudaGetLastError());                                          \
    } while (0)

template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    // grow-only (re)allocation; contents are not preserved
    void reserve(size_t count) {
        if (count <= n) return;
        release();
        CSG_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        n = count;
    }
};

inline unsigned ilog2(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) l++; return l; }

}  // namespace csg
