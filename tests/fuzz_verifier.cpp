// Mutation fuzzer for the product's verifier (csrc/host/verifier.cpp), built with -fsanitize=address,undefined by
// tests/test_host.py::test_verifier_survives_mutated_proofs.  A proof is attacker-controlled input (ADVICE round 1 found an
// out-of-bounds access behind a crafted context): every mutation of a valid proof must come back as a rejection code -- not as a
// crash, not as a sanitizer report, and not as CSG_OK.
//
//   fuzz_verifier <air id> <pub.bin (u64 words)> <proof.bin> <iterations> <seed>
// prints "mutations N accepted A codes {...}" and exits 0 when nothing was accepted; the sanitizers abort the process otherwise.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../include/csg.h"

static std::vector<uint8_t> slurp(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    std::vector<uint8_t> v;
    uint8_t buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + k);
    fclose(f);
    return v;
}

static uint64_t rng_state;
static uint64_t rnd() {   // splitmix64
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

int main(int argc, char **argv) {
    if (argc != 6) { fprintf(stderr, "usage: fuzz_verifier air pub.bin proof.bin iterations seed\n"); return 2; }
    const int air = atoi(argv[1]);
    const std::vector<uint8_t> pubb = slurp(argv[2]), proof = slurp(argv[3]);
    const long iters = atol(argv[4]);
    rng_state = strtoull(argv[5], nullptr, 10);
    std::vector<uint64_t> pub(pubb.size() / 8);
    memcpy(pub.data(), pubb.data(), pub.size() * 8);
    if (csg_verify(air, pub.data(), pub.size(), proof.data(), proof.size()) != CSG_OK) { fprintf(stderr, "the unmodified proof does not verify\n"); return 3; }

    std::map<int, long> codes;
    long accepted = 0, done = 0;
    for (long it = 0; it < iters; it++) {
        std::vector<uint8_t> p = proof;
        const unsigned kind = (unsigned)(rnd() % 8);
        // the first ~40 bytes are the context (widths, log sizes, options): half of the single-byte mutations aim there
        auto pos = [&]() -> size_t { return (rnd() & 1) ? (size_t)(rnd() % (p.size() < 40 ? p.size() : 40)) : (size_t)(rnd() % p.size()); };
        switch (kind) {
        case 0: p[pos()] ^= (uint8_t)(1u << (rnd() % 8)); break;                                        // one bit
        case 1: p[pos()] = (uint8_t)rnd(); break;                                                      // one byte
        case 2: { size_t a = pos(); p[a] = 0xff; if (a + 1 < p.size()) p[a + 1] = 0xff; break; }        // a length field blown up
        case 3: p.resize((size_t)(rnd() % p.size())); break;                                           // truncated
        case 4: { size_t a = (size_t)(rnd() % p.size()), n = 1 + (size_t)(rnd() % 64); p.insert(p.begin() + a, n, (uint8_t)rnd()); break; }   // bytes inserted
        case 5: { size_t a = (size_t)(rnd() % p.size()), n = 1 + (size_t)(rnd() % 64); if (a + n > p.size()) n = p.size() - a; p.erase(p.begin() + a, p.begin() + a + n); break; }   // bytes removed
        case 6: { size_t a = (size_t)(rnd() % p.size()), b = (size_t)(rnd() % p.size()), n = 1 + (size_t)(rnd() % 256);                              // a chunk copied elsewhere
                  for (size_t k = 0; k < n && a + k < p.size() && b + k < p.size(); k++) p[a + k] = proof[b + k]; break; }
        default: { size_t a = (size_t)(rnd() % p.size()), n = 1 + (size_t)(rnd() % 32); for (size_t k = 0; k < n && a + k < p.size(); k++) p[a + k] = 0; break; }   // zeroed run
        }
        if (p == proof) continue;
        done++;
        // a fresh exact-size heap copy, so that a read past the end is a read past an allocation
        uint8_t *heap = p.empty() ? nullptr : (uint8_t *)malloc(p.size());
        if (!p.empty()) memcpy(heap, p.data(), p.size());
        const int rc = heap ? csg_verify(air, pub.data(), pub.size(), heap, p.size()) : csg_verify(air, pub.data(), pub.size(), (const uint8_t *)"", 0);
        free(heap);
        codes[rc]++;
        if (rc == CSG_OK) {
            accepted++;
            fprintf(stderr, "ACCEPTED a mutated proof: iteration %ld kind %u length %zu (original %zu)\n", it, kind, p.size(), proof.size());
        }
    }
    printf("mutations %ld accepted %ld codes {", done, accepted);
    for (auto &kv : codes) printf(" %d: %ld", kv.first, kv.second);
    printf(" }\n");
    return accepted ? 1 : 0;
}
