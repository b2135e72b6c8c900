// Test harness (CPU only): checks the host-compilable halves of the product against the oracle.
//   1. air_desc.cpp  (degrees, periodic columns, assertions, ce blowup)          vs  oracle/airs.c air_new()
//   2. airs.cuh      (fused per-row transition evaluation + linear combination)    vs  oracle air->eval() + explicit sum
//   3. transcript.hpp batch-opening shape                                          vs  oracle merkle_prove_batch()
//   4. with a file argument (tests/golden/air_vectors.txt): BOTH of the above sides against the golden vectors of the
//      independent Python restatement (oracle/pyair.py): degrees, periodic columns, assertions, full result[] vectors
// The same airs.cuh code is what the CUDA kernels run per row; here it is compiled by g++.
// Build: g++ -O2 -std=c++17 tests/host_harness.cpp certificate_stark_b200/csrc/host/air_desc.cpp -Loracle -loracle
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
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

// ---------------------------------------------------------------------------------------------- golden vectors (oracle/pyair.py)
extern "C" void sha3_256(const uint8_t *in, size_t len, uint8_t out[32]);
static std::vector<uint64_t> hex_words(std::istringstream &in) {
    std::vector<uint64_t> v;
    std::string t;
    while (in >> t) v.push_back(strtoull(t.c_str(), nullptr, 16));
    return v;
}
static std::vector<fe> to_mont_vec(const std::vector<uint64_t> &v) { std::vector<fe> r; for (uint64_t x : v) r.push_back(f63::to_mont(x)); return r; }
static std::string column_fingerprint(const fe *col, size_t len) {
    std::vector<uint8_t> bytes(8 * len);
    for (size_t i = 0; i < len; i++) { uint64_t c = f63::from_mont(col[i]); memcpy(&bytes[8 * i], &c, 8); }
    uint8_t dg[32];
    sha3_256(bytes.data(), bytes.size(), dg);
    char hex[65];
    for (int i = 0; i < 32; i++) snprintf(hex + 2 * i, 3, "%02x", dg[i]);
    return hex;
}
template <int AIR>
static fe product_merged(const csg::AirDesc &d, const csg::TransitionGroups &tg, const std::vector<fe> &cur, const std::vector<fe> &next, std::vector<fe> pv,
                         const std::vector<fe> &alpha, const std::vector<fe> &beta, const std::vector<fe> &xp) {
    const size_t w = d.width, np = pv.size();
    pv.push_back(0);
    std::vector<fe> m(2 * w);
    for (size_t c = 0; c < w; c++) { m[2 * c] = cur[c]; m[2 * c + 1] = next[c]; }
    std::vector<uint32_t> off(np + 1), mask(np + 1, 0);
    for (size_t c = 0; c <= np; c++) off[c] = (uint32_t)c;
    airs::Frame f{m.data(), m.data() + 1, 2};
    airs::Periodic P{pv.data(), off.data(), mask.data(), 0};
    airs::Comb C{alpha.data(), beta.data(), tg.group_of.data(), xp.data(), 1, f63::acc192(), nullptr, 0};
    airs::eval_transition<AIR>(f, P, C);
    return C.sum.reduce();
}
static int check_golden(const char *path) {
    std::ifstream file(path);
    if (!file) { printf("cannot open %s\n", path); return 1; }
    std::string line;
    size_t airs_seen = 0, frames_seen = 0;
    while (std::getline(file, line)) {
        if (line.empty() || line[0] == '#') continue;
        if (line == "END") break;
        std::istringstream hdr(line);
        std::string tag, name, k;
        int id; size_t n, npub, width, nc, np, na, nrows, ce;
        hdr >> tag >> id >> name >> k >> n >> k >> npub >> k >> width >> k >> nc >> k >> np >> k >> na >> k >> nrows >> k >> ce;
        if (tag != "AIR") { printf("bad golden file at: %s\n", line.c_str()); return 1; }
        airs_seen++;
        std::getline(file, line);
        std::istringstream pl(line.substr(4));
        std::vector<uint64_t> pub = hex_words(pl);
        CHECK(pub.size() == npub, "golden %s: public input count", name.c_str());
        csg::AirDesc d = csg::make_air(id, n, pub.data(), pub.size());
        air_t *o = air_new(id, n, pub.data(), pub.size());
        CHECK(o != nullptr, "golden %s: oracle air_new failed", name.c_str());
        CHECK(d.width == width && o->width == width && d.num_constraints() == nc && o->num_constraints == nc, "golden %s: shape", name.c_str());
        CHECK(d.ce_blowup() == ce && air_ce_blowup(o) == ce, "golden %s: ce blowup", name.c_str());
        for (size_t i = 0; i < nc; i++) {          // declared degrees, constraint by constraint
            std::getline(file, line);
            std::istringstream dl(line.substr(4));
            uint32_t base; size_t ncyc; dl >> base >> ncyc;
            std::vector<uint32_t> cyc(ncyc); for (auto &c : cyc) dl >> c;
            bool pd = d.degrees[i].base == base && d.degrees[i].cycles == cyc;
            bool od = o->degrees[i].base == base && o->degrees[i].ncycles == ncyc;
            for (size_t c = 0; od && c < ncyc; c++) od = o->degrees[i].cycles[c] == cyc[c];
            CHECK(pd, "golden %s: product degree of constraint %zu", name.c_str(), i);
            CHECK(od, "golden %s: oracle degree of constraint %zu", name.c_str(), i);
        }
        CHECK(d.periodic.size() == np && o->num_periodic == np, "golden %s: periodic column count", name.c_str());
        for (size_t c = 0; c < np; c++) {          // periodic columns by length and SHA3-256 of their canonical words
            std::getline(file, line);
            std::istringstream pc(line.substr(4));
            size_t len; std::string fp; pc >> len >> fp;
            CHECK(d.periodic[c].values.size() == len && column_fingerprint(d.periodic[c].values.data(), len) == fp, "golden %s: product periodic column %zu", name.c_str(), c);
            CHECK(o->periodic_len[c] == len && column_fingerprint(o->periodic[c], len) == fp, "golden %s: oracle periodic column %zu", name.c_str(), c);
        }
        CHECK(d.assertions.size() == na && o->num_assertions == na, "golden %s: assertion count", name.c_str());
        for (size_t i = 0; i < na; i++) {          // assertions as a multiset (winterfell sorts them; the golden file has get_assertions() order)
            std::getline(file, line);
            std::istringstream al(line.substr(4));
            std::string kind; uint32_t col; size_t first, stride, nv; al >> kind >> col >> first >> stride >> nv;
            std::vector<fe> vals = to_mont_vec(hex_words(al));
            if (kind == "sequence" && nv == 1) stride = 0;   // winterfell: a sequence of one value is a single assertion [RECALLED]
            bool pf = false, of = false;
            for (const csg::Assertion &a : d.assertions) pf = pf || (a.column == col && a.first_step == first && a.stride == stride && a.values == vals);
            for (uint32_t q = 0; q < o->num_assertions; q++) {
                const air_assertion &b = o->assertions[q];
                of = of || (b.column == col && b.first_step == first && b.stride == stride && b.nvalues == nv && !memcmp(b.values, vals.data(), nv * sizeof(fe)));
            }
            CHECK(pf, "golden %s: product has no assertion (col %u, step %zu, stride %zu)", name.c_str(), col, first, stride);
            CHECK(of, "golden %s: oracle has no assertion (col %u, step %zu, stride %zu)", name.c_str(), col, first, stride);
        }
        csg::TransitionGroups tg = csg::transition_groups(d);
        for (size_t r = 0; r < nrows; r++) {
            std::string rl, cl, nl, pl2, sl;
            std::getline(file, rl); std::getline(file, cl); std::getline(file, nl); std::getline(file, pl2); std::getline(file, sl);
            long step = atol(rl.c_str() + 4);
            std::istringstream ci(cl.substr(4)), ni(nl.substr(4)), pi(pl2.size() > 3 ? pl2.substr(3) : ""), si(sl.substr(4));
            std::vector<fe> cur = to_mont_vec(hex_words(ci)), next = to_mont_vec(hex_words(ni)), pv = to_mont_vec(hex_words(pi)), want = to_mont_vec(hex_words(si));
            CHECK(cur.size() == width && next.size() == width && pv.size() == np && want.size() == nc, "golden %s frame %zu: sizes", name.c_str(), r);
            frames_seen++;
            // (a) the C oracle, slot by slot
            std::vector<fe> res(nc, 0), pvo(pv);
            pvo.push_back(0);
            o->eval(o, cur.data(), next.data(), pvo.data(), res.data());
            for (size_t i = 0; i < nc; i++) CHECK(res[i] == want[i], "golden %s frame %zu (step %ld): oracle/airs.c result[%zu] differs", name.c_str(), r, step, i);
            if (step >= 0) {   // witness rows: the oracle's own periodic tables at that step must give the same (vanishing) values
                std::vector<fe> res2(nc, 0);
                air_eval_row(o, (size_t)step, cur.data(), next.data(), res2.data());
                for (size_t i = 0; i < nc; i++) CHECK(res2[i] == want[i], "golden %s step %ld: oracle periodic values, result[%zu]", name.c_str(), step, i);
            }
            // (b) the product never materialises result[]: compare its merged value with the same random combination of the golden slots,
            // for two independent draws of the coefficients
            for (int rep = 0; rep < 2; rep++) {
                std::vector<fe> alpha(nc), beta(nc), xp(tg.adj.size());
                for (auto &v : alpha) v = rnd();
                for (auto &v : beta) v = rnd();
                for (auto &v : xp) v = rnd();
                fe expect = 0;
                for (size_t i = 0; i < nc; i++) expect = f63::add(expect, f63::mul(want[i], f63::add(alpha[i], f63::mul(beta[i], xp[tg.group_of[i]]))));
                fe got = id == 0 ? product_merged<0>(d, tg, cur, next, pv, alpha, beta, xp) : id == 1 ? product_merged<1>(d, tg, cur, next, pv, alpha, beta, xp)
                       : id == 2 ? product_merged<2>(d, tg, cur, next, pv, alpha, beta, xp) : id == 3 ? product_merged<3>(d, tg, cur, next, pv, alpha, beta, xp)
                       : id == 4 ? product_merged<4>(d, tg, cur, next, pv, alpha, beta, xp) : product_merged<5>(d, tg, cur, next, pv, alpha, beta, xp);
                CHECK(got == expect, "golden %s frame %zu (step %ld): product airs.cuh merged value differs", name.c_str(), r, step);
            }
        }
        air_free(o);
    }
    if (failures) printf("%d FAILURES\n", failures);
    else printf("golden vectors: %zu AIR sections, %zu frames, all checks passed\n", airs_seen, frames_seen);
    return failures ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc > 1) return check_golden(argv[1]);
    // the field self-check the device runs (csg_debug_field_selftest), on the host forms of the same arithmetic: pins the
    // checker, so that a mismatch on the device is a device bug
    for (uint64_t seed = 0; seed < 512; seed++) {
        unsigned long long bad = f63::field_selfcheck(0x5eedULL + seed * 0x9e3779b97f4a7c15ULL, 64);
        CHECK(bad == 0, "field self-check: %llu mismatches for seed %llu", bad, (unsigned long long)seed);
    }
    // the dedicated Fp6 squaring (21 multiply-accumulates) against the general product, edge values included
    for (int it = 0; it < 20000; it++) {
        ecc::fp6 a;
        for (int i = 0; i < 6; i++) a.c[i] = it < 128 ? (((it >> i) & 1) ? f63::P - 1 : ((it & 64) ? 1 : 0)) : rnd();
        const ecc::fp6 sq = ecc::sqr(a), mm = ecc::mul(a, a);
        bool same = true;
        for (int i = 0; i < 6; i++) same = same && sq.c[i] == mm.c[i];
        CHECK(same, "fp6 squaring differs from the product at iteration %d", it);
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
