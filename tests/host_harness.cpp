// Test harness (CPU only): checks the host-compilable halves of the product against the oracle.
//   1. air_desc.cpp  (degrees, periodic columns, assertions, ce blowup)          vs  oracle/airs.c air_new()
//   2. airs.cuh      (fused per-row transition evaluation + linear combination)    vs  oracle air->eval() + explicit sum
//   3. transcript.hpp batch-opening shape                                          vs  oracle merkle_prove_batch()
// The same airs.cuh code is what the CUDA kernels run per row; here it is compiled by g++.
// Build: g++ -O2 -std=c++17 tests/host_harness.cpp certificate_stark_b200/csrc/host/air_desc.cpp -Loracle -loracle
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../certificate_stark_b200/csrc/airs.cuh"
#include "../certificate_stark_b200/csrc/field_selfcheck.cuh"
#include "../certificate_stark_b200/csrc/host/air_desc.hpp"
#include "../certificate_stark_b200/csrc/host/transcript.hpp"

extern "C" {
#include "../oracle/air.h"
#include "../oracle/stark.h"
}

using f63::fe;
static std::mt19937_64 rng(12345);
static fe rnd() { return rng() % f63::P; }
static int failures = 0;
#define CHECK(c, ...) do { if (!(c)) { failures++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

static std::vector<uint64_t> make_pub(int id, size_t n) {
    size_t np = id == 2 ? 29 : id == 3 ? 38 * (n / 512) : id == 4 ? 1 : 14;
    std::vector<uint64_t> p(np);
    for (auto &v : p) v = rng() % f63::P;
    return p;
}

template <int AIR>
static void check_eval(const csg::AirDesc &d, air_t *o) {
    const size_t w = d.width, nc = d.num_constraints(), np = d.periodic.size();
    csg::TransitionGroups tg = csg::transition_groups(d);
    for (int rep = 0; rep < 20; rep++) {
        std::vector<fe> cur(w), next(w), pv(np + 1), alpha(nc), beta(nc), xp(tg.adj.size());
        for (auto &v : cur) v = rnd();
        for (auto &v : next) v = rnd();
        for (auto &v : pv) v = rnd();
        for (auto &v : alpha) v = rnd();
        for (auto &v : beta) v = rnd();
        for (auto &v : xp) v = rnd();
        if (rep < 4) for (size_t i = 0; i < w; i++) { cur[i] = (rep & 1) ? f63::ONE : 0; next[i] = (rep & 2) ? f63::ONE : 0; }   // 0/1 corner cases
        std::vector<fe> res(nc, 0);
        o->eval(o, cur.data(), next.data(), pv.data(), res.data());
        fe expect = 0;
        for (size_t i = 0; i < nc; i++) expect = f63::add(expect, f63::mul(res[i], f63::add(alpha[i], f63::mul(beta[i], xp[tg.group_of[i]]))));
        // frame laid out as two rows of a column-major matrix with 2 rows
        std::vector<fe> m(2 * w);
        for (size_t c = 0; c < w; c++) { m[2 * c] = cur[c]; m[2 * c + 1] = next[c]; }
        std::vector<uint32_t> off(np + 1), mask(np + 1, 0);
        for (size_t c = 0; c <= np; c++) off[c] = (uint32_t)c;
        airs::Frame f{m.data(), m.data() + 1, 2};
        airs::Periodic P{pv.data(), off.data(), mask.data(), 0};
        airs::Comb C{alpha.data(), beta.data(), tg.group_of.data(), xp.data(), 1, f63::acc192(), nullptr, 0};
        airs::eval_transition<AIR>(f, P, C);
        fe got = C.sum.reduce();
        CHECK(got == expect, "air %d rep %d: fused evaluation %016llx != oracle %016llx", AIR, rep, (unsigned long long)got, (unsigned long long)expect);
        // split mode: T = A + sum_g x^adj_g * B_g with the alpha and per-group beta parts accumulated separately
        std::vector<uint64_t> parts(3 * airs::MAX_SPLIT_GROUPS, 0);
        airs::SplitComb S{alpha.data(), beta.data(), tg.group_of.data(), nullptr, 0, f63::acc192(), parts.data(), 1};
        airs::eval_transition<AIR>(f, P, S);
        fe t = S.sum.reduce();
        for (size_t g = 0; g < tg.adj.size(); g++) t = f63::add(t, f63::mul(xp[g], S.part_value((int)g)));
        CHECK(t == expect, "air %d rep %d: split evaluation differs from the oracle", AIR, rep);
        // the device form of the Rescue constraints: forward MDS product folded into per-proof coefficient tables
        // (rescue_tables.h; on the host only when tables are supplied), combined and split
        airs::RescueTables RT;
        airs::fill_rescue_tables(AIR, alpha.data(), beta.data(), tg.group_of.data(), (unsigned)nc, RT);
        airs::Comb C2{alpha.data(), beta.data(), tg.group_of.data(), xp.data(), 1, f63::acc192(), nullptr, 0};
        C2.rt = &RT;
        airs::eval_transition<AIR>(f, P, C2);
        CHECK(C2.sum.reduce() == expect, "air %d rep %d: evaluation through the Rescue tables differs from the oracle", AIR, rep);
        std::vector<uint64_t> parts2(3 * airs::MAX_SPLIT_GROUPS, 0);
        airs::SplitComb S2{alpha.data(), beta.data(), tg.group_of.data(), nullptr, 0, f63::acc192(), parts2.data(), 1};
        S2.rt = &RT;
        airs::eval_transition<AIR>(f, P, S2);
        fe t2 = S2.sum.reduce();
        for (size_t g = 0; g < tg.adj.size(); g++) t2 = f63::add(t2, f63::mul(xp[g], S2.part_value((int)g)));
        CHECK(t2 == expect, "air %d rep %d: split evaluation through the Rescue tables differs from the oracle", AIR, rep);
    }
}

static void check_air(int id, size_t n) {
    std::vector<uint64_t> pub = make_pub(id, n);
    csg::AirDesc d = csg::make_air(id, n, pub.data(), pub.size());
    air_t *o = air_new(id, n, pub.data(), pub.size());
    CHECK(o != nullptr, "oracle air_new failed");
    CHECK(d.width == o->width && d.num_constraints() == o->num_constraints, "air %d: shape", id);
    CHECK(d.ce_blowup() == air_ce_blowup(o), "air %d: ce blowup %zu vs %zu", id, d.ce_blowup(), air_ce_blowup(o));
    for (size_t i = 0; i < d.num_constraints(); i++) {
        const air_degree &od = o->degrees[i];
        bool same = d.degrees[i].base == od.base && d.degrees[i].cycles.size() == od.ncycles;
        for (size_t k = 0; same && k < od.ncycles; k++) same = d.degrees[i].cycles[k] == od.cycles[k];
        CHECK(same, "air %d: degree of constraint %zu", id, i);
        CHECK(d.evaluation_degree(d.degrees[i]) == air_eval_degree(&od, n), "air %d: evaluation degree %zu", id, i);
    }
    CHECK(d.periodic.size() == o->num_periodic, "air %d: %zu periodic columns vs %u", id, d.periodic.size(), o->num_periodic);
    for (size_t c = 0; c < d.periodic.size() && c < o->num_periodic; c++) {
        bool same = d.periodic[c].values.size() == o->periodic_len[c];
        for (size_t i = 0; same && i < o->periodic_len[c]; i++) same = d.periodic[c].values[i] == o->periodic[c][i];
        CHECK(same, "air %d: periodic column %zu", id, c);
    }
    CHECK(d.assertions.size() == o->num_assertions, "air %d: assertion count", id);
    // same multiset of assertions, and sorted by (stride, first_step, column)
    for (size_t i = 0; i < d.assertions.size(); i++) {
        const csg::Assertion &a = d.assertions[i];
        bool found = false;
        for (uint32_t k = 0; k < o->num_assertions && !found; k++) {
            const air_assertion &b = o->assertions[k];
            found = a.column == b.column && a.first_step == b.first_step && a.stride == b.stride && a.values.size() == b.nvalues &&
                    !memcmp(a.values.data(), b.values, b.nvalues * sizeof(fe));
        }
        CHECK(found, "air %d: assertion %zu has no counterpart", id, i);
        if (i) {
            const csg::Assertion &p = d.assertions[i - 1];
            bool ordered = p.stride < a.stride || (p.stride == a.stride && (p.first_step < a.first_step || (p.first_step == a.first_step && p.column <= a.column)));
            CHECK(ordered, "air %d: assertion order at %zu", id, i);
        }
    }
    switch (id) {
    case 0: check_eval<0>(d, o); break;
    case 1: check_eval<1>(d, o); break;
    case 2: check_eval<2>(d, o); break;
    case 3: check_eval<3>(d, o); break;
    case 4: check_eval<4>(d, o); break;
    default: check_eval<5>(d, o); break;
    }
    air_free(o);
}

static void check_batch_openings() {
    for (int rep = 0; rep < 50; rep++) {
        const size_t nl = (size_t)1 << (2 + rng() % 10);
        std::vector<uint8_t> nodes(2 * nl * 32);
        for (auto &b : nodes) b = (uint8_t)rng();
        size_t np = 1 + rng() % std::min<size_t>(nl - 1, 42);
        std::vector<size_t> pos;
        while (pos.size() < np) { size_t p = rng() % nl; if (std::find(pos.begin(), pos.end(), p) == pos.end()) pos.push_back(p); }
        std::vector<uint8_t> expect(1 + np * (1 + 32 * 24));
        size_t elen = merkle_prove_batch(nodes.data(), nl, pos.data(), np, expect.data());
        auto slots = csg::batch_opening_nodes(nl, pos);
        std::vector<uint8_t> got;
        got.push_back((uint8_t)slots.size());
        for (auto &s : slots) { got.push_back((uint8_t)s.size()); for (uint32_t i : s) got.insert(got.end(), nodes.begin() + 32 * (size_t)i, nodes.begin() + 32 * (size_t)i + 32); }
        CHECK(got.size() == elen && !memcmp(got.data(), expect.data(), elen), "batch opening shape, %zu leaves %zu positions", nl, np);
    }
}

int main() {
    // the field self-check the device runs (csg_debug_field_selftest), on the host forms of the same arithmetic: pins the
    // checker, so that a mismatch on the device is a device bug
    for (uint64_t seed = 0; seed < 512; seed++) {
        unsigned long long bad = f63::field_selfcheck(0x5eedULL + seed * 0x9e3779b97f4a7c15ULL, 64);
        CHECK(bad == 0, "field self-check: %llu mismatches for seed %llu", bad, (unsigned long long)seed);
    }
    check_air(0, 2048);
    check_air(1, 1024);
    check_air(2, 16);
    check_air(3, 512);
    check_air(3, 2048);
    check_air(4, 64);
    check_air(5, 64);
    check_batch_openings();
    printf(failures ? "%d FAILURES\n" : "host harness: all checks passed\n", failures);
    return failures ? 1 : 0;
}
