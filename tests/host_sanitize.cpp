// Host witness and batch builders (csrc/host/witness.cpp) driven through the C ABI with exact-size heap buffers; built with
// AddressSanitizer + UBSan + LeakSanitizer by tests/test_host.py::test_host_builders_are_sanitizer_clean.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/csg.h"
// exact-size heap buffers: an overrun of any builder is an overrun of an allocation
int main() {
    int bad = 0;
    for (unsigned depth : {3u, 15u}) for (size_t ntx : {1, 2, 8}) {
        csg_tx_batch *b = csg_tx_batch_new(5 + depth, ntx, depth);
        if (!b) { printf("batch %zu %u failed\n", ntx, depth); return 1; }
        uint64_t *tr = (uint64_t *)malloc(94 * 1024 * ntx * 8), pub[14], r0[7], r1[7];
        bad += csg_build_trace_transaction(b, tr, pub) != 0;
        free(tr);
        tr = (uint64_t *)malloc(65 * 512 * ntx * 8);
        bad += csg_build_trace_merkle_update(b, tr, pub) != 0;
        free(tr);
        csg_tx_batch_roots(b, r0, r1);
        size_t w = csg_tx_batch_pack(b, nullptr);
        uint64_t *rec = (uint64_t *)malloc(w * 8);
        csg_tx_batch_pack(b, rec);
        free(rec);
        csg_tx_batch_free(b);
    }
    for (size_t ns : {1, 2, 4}) {
        csg_sig_batch *s = csg_sig_batch_new(9, ns);
        uint64_t *tr = (uint64_t *)malloc(56 * 512 * ns * 8), *pub = (uint64_t *)malloc(38 * ns * 8);
        bad += csg_build_trace_schnorr(s, tr, pub) != 0;
        size_t w = csg_sig_batch_pack(s, nullptr);
        uint64_t *rec = (uint64_t *)malloc(w * 8);
        csg_sig_batch_pack(s, rec);
        free(rec); free(tr); free(pub);
        csg_sig_batch_free(s);
    }
    {
        uint64_t *tr = (uint64_t *)malloc(2 * 64 * 8), pub[1];
        bad += csg_build_trace_range(0x7fffffffffffffffULL, tr, pub) != 0;
        bad += csg_build_trace_range(0, tr, pub) != 0;
        free(tr);
        uint64_t seed[7] = {42, 43, 44, 45, 46, 47, 48}, pubr[14];
        for (size_t len : {1, 16, 128}) { tr = (uint64_t *)malloc(14 * 8 * len * 8); bad += csg_build_trace_rescue(seed, len, tr, pubr) != 0; free(tr); }
    }
    printf("witness builders under sanitizers: %d failures\n", bad);
    return bad;
}
