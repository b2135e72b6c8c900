// See air_desc.hpp.  Masks are written as predicates on the step inside a cycle (the reference stitches vectors with
// utils/periodic_columns.rs; the resulting columns are the same: SURVEY.md Appendix B lists them).
#include "air_desc.hpp"

#include <algorithm>
#include <functional>
#include <stdexcept>

#include "../ref_constants.h"

namespace csg {
using namespace f63;

namespace {
enum : size_t {
    HSW = 14, HRW = 7, APW = 12, PPW = 18, PCW = 6,
    SENDER_INITIAL = 0, SENDER_UPDATED = 15, RECEIVER_INITIAL = 29, RECEIVER_BIT = 43, RECEIVER_UPDATED = 44, PREV_ROOT = 58,
    MERKLE_WIDTH = 65, MERKLE_CONSTRAINTS = 106, INT_ROOT_RES = 92, TX_WIDTH = 94, TX_CONSTRAINTS = 115, SCHNORR_WIDTH = 56,
    MERKLE_CYCLE = 512, TX_CYCLE = 1024, SIG_CYCLE = 512, SCALAR_MUL_LENGTH = 510, NUM_HASH_ITER = 5, RANGE_LOG = 64,
    TREE_DEPTH = 15, HASH_LEN = 8 * TREE_DEPTH + 7,   // src/merkle/constants.rs:21-31 (non-test build)
    SIG_PUB_WORDS = 38                                // message[28] Rx[6] s[4]
};

PeriodicColumn mask(size_t period, const std::function<bool(size_t)> &on) {
    PeriodicColumn c;
    c.values.resize(period);
    for (size_t i = 0; i < period; i++) c.values[i] = on(i) ? ONE : ZERO;
    return c;
}
// the 28 round-constant columns, cycle 8 (src/utils/rescue.rs:303-318)
void push_ark(std::vector<PeriodicColumn> &cols) {
    for (size_t j = 0; j < 2 * HSW; j++) {
        PeriodicColumn c;
        for (size_t i = 0; i < 8; i++) c.values.push_back(CSG_ARK_M[i * 28 + j]);
        cols.push_back(c);
    }
}
Degree deg(uint32_t base, std::initializer_list<uint32_t> cycles = {}) { return Degree{base, cycles}; }

// src/merkle/update/air.rs:371-401
std::vector<Degree> merkle_update_degrees(uint32_t cycle) {
    std::vector<Degree> d;
    for (int half = 0; half < 2; half++) {
        for (size_t i = 0; i < HSW; i++) d.push_back(deg(3, {cycle}));
        d.push_back(deg(2, {cycle}));
        for (size_t i = 0; i < HSW; i++) d.push_back(deg(3, {cycle}));
    }
    while (d.size() < MERKLE_CONSTRAINTS) d.push_back(deg(1, {cycle}));
    return d;
}
// src/schnorr/air.rs:533-585
std::vector<Degree> schnorr_degrees(size_t num_signatures, uint32_t cycle) {
    const uint32_t bit_degree = num_signatures == 1 ? 3 : 5;
    std::vector<Degree> d;
    for (size_t i = 0; i < PCW; i++) d.push_back(deg(5, {cycle, cycle}));
    for (size_t i = 0; i < APW; i++) d.push_back(deg(4, {cycle, cycle}));
    d.push_back(deg(2, {cycle}));
    for (size_t i = 0; i < PPW; i++) d.push_back(deg(bit_degree, {cycle, cycle}));
    d.push_back(deg(2, {cycle}));
    for (size_t i = 0; i < 4; i++) d.push_back(deg(1, {cycle, cycle}));
    for (size_t i = 0; i < HSW; i++) d.push_back(deg(3, {cycle}));
    return d;
}
// the eight mask columns shared by SchnorrAir and the Schnorr half of the transaction AIR; s = step inside the signature
// (src/schnorr/air.rs:334-391): global, scalar mult, doubling, digest limb 0..3, hash
std::vector<std::function<bool(size_t)>> schnorr_predicates() {
    return {
        [](size_t s) { return s < SCALAR_MUL_LENGTH + 1; },
        [](size_t s) { return s < SCALAR_MUL_LENGTH; },
        [](size_t s) { return s < SCALAR_MUL_LENGTH && s % 2 == 0; },
        [](size_t s) { return s < 126; },
        [](size_t s) { return s >= 126 && s < 254; },
        [](size_t s) { return s >= 254 && s < 382; },
        [](size_t s) { return s >= 382 && s < 510; },
        [](size_t s) { return s < 8 * NUM_HASH_ITER && s % 8 < 7; },
    };
}
Assertion single(size_t col, size_t step, fe v) { return Assertion{(uint32_t)col, step, 0, {v}}; }
Assertion periodic(size_t col, size_t first, size_t stride, fe v) { return Assertion{(uint32_t)col, first, stride, {v}}; }
// a one-value sequence is a single assertion
Assertion sequence(size_t col, size_t first, size_t stride, const std::vector<fe> &v) { return Assertion{(uint32_t)col, first, v.size() == 1 ? 0 : stride, v}; }
fe canon(uint64_t v) { return to_mont(v % P); }

void need(bool ok, const char *what) { if (!ok) throw std::invalid_argument(what); }
}  // namespace

size_t AirDesc::ce_blowup() const {
    size_t m = 2;
    for (const Degree &d : degrees) {
        size_t v = d.base + d.cycles.size(), p2 = 1;
        while (p2 < v) p2 *= 2;
        m = std::max(m, p2);
    }
    return m;
}
size_t AirDesc::evaluation_degree(const Degree &d) const {
    size_t r = (size_t)d.base * (trace_len - 1);
    for (uint32_t c : d.cycles) r += (trace_len / c) * (c - 1);
    return r;
}

AirDesc make_air(int air_id, size_t n, const uint64_t *pub, size_t npub) {
    need(n >= 8 && (n & (n - 1)) == 0, "trace length must be a power of two, at least 8");
    AirDesc a;
    a.id = air_id; a.trace_len = n;
    a.pub_inputs.assign(pub, pub + npub);
    switch (air_id) {
    case 0: {   // TransactionAir
        need(npub == 14 && n % TX_CYCLE == 0, "transaction AIR: 14 public inputs, trace length a multiple of 1024");
        a.width = TX_WIDTH;
        a.degrees = merkle_update_degrees(TX_CYCLE);
        a.degrees[RECEIVER_BIT] = deg(3, {TX_CYCLE});
        a.degrees[INT_ROOT_RES] = deg(2, {TX_CYCLE});
        std::vector<Degree> sd = schnorr_degrees(2, TX_CYCLE);
        for (size_t i = 0; i < PPW; i++) { a.degrees[i] = sd[i]; a.degrees[i + PPW + 1] = sd[i + PPW + 1]; }
        while (a.degrees.size() < TX_CONSTRAINTS) a.degrees.push_back(deg(1, {TX_CYCLE}));
        // masks over the 1024-step transaction: Merkle phase in steps 0..511, signature + range proofs in 512..1023
        auto in_sig = [](size_t i) { return i >= MERKLE_CYCLE; };
        a.periodic.push_back(mask(TX_CYCLE, [](size_t i) { return i == 0; }));                          // setup
        a.periodic.push_back(mask(TX_CYCLE, [](size_t i) { return i < HASH_LEN; }));                    // merkle
        a.periodic.push_back(mask(8, [](size_t i) { return i == 7; }));                                 // hash input
        a.periodic.push_back(mask(TX_CYCLE, [](size_t i) { return i == HASH_LEN - 1; }));               // finish
        a.periodic.push_back(mask(TX_CYCLE, [](size_t i) { return i < HASH_LEN && i % 8 < 7; }));       // hash
        for (auto &pred : schnorr_predicates()) {
            if (a.periodic.size() == 12) break;   // global, scalar mult, doubling, digest x4 -> columns 5..11
            a.periodic.push_back(mask(TX_CYCLE, [&](size_t i) { return in_sig(i) && pred(i - MERKLE_CYCLE); }));
        }
        { auto hp = schnorr_predicates()[7]; a.periodic.push_back(mask(TX_CYCLE, [&](size_t i) { return in_sig(i) && hp(i - MERKLE_CYCLE); })); }   // 12
        for (size_t k = 0; k < NUM_HASH_ITER - 1; k++)                                                   // 13..16
            a.periodic.push_back(mask(TX_CYCLE, [&](size_t i) { return in_sig(i) && i - MERKLE_CYCLE == (k + 1) * 8 - 1; }));
        a.periodic.push_back(mask(TX_CYCLE, [&](size_t i) { return in_sig(i) && i - MERKLE_CYCLE < RANGE_LOG; }));        // range step
        a.periodic.push_back(mask(TX_CYCLE, [&](size_t i) { return in_sig(i) && i - MERKLE_CYCLE == RANGE_LOG - 1; }));   // range finish
        a.periodic.push_back(mask(TX_CYCLE, [](size_t i) { return i >= 1 && i < MERKLE_CYCLE + RANGE_LOG; }));           // value copy
        push_ark(a.periodic);
        a.assertions = {single(PREV_ROOT, 0, canon(pub[0])), single(PREV_ROOT + 1, 0, canon(pub[1])),
                        single(PREV_ROOT, n - 1, canon(pub[7])), single(PREV_ROOT + 1, n - 1, canon(pub[8]))};
        break;
    }
    case 1: {   // MerkleAir
        need(npub == 14 && n % MERKLE_CYCLE == 0, "merkle-update AIR: 14 public inputs, trace length a multiple of 512");
        a.width = MERKLE_WIDTH;
        a.degrees = merkle_update_degrees(MERKLE_CYCLE);
        a.periodic.push_back(mask(MERKLE_CYCLE, [](size_t i) { return i == 0; }));
        a.periodic.push_back(mask(MERKLE_CYCLE, [](size_t i) { return i < HASH_LEN; }));
        a.periodic.push_back(mask(8, [](size_t i) { return i == 7; }));
        a.periodic.push_back(mask(MERKLE_CYCLE, [](size_t i) { return i == HASH_LEN - 1; }));
        a.periodic.push_back(mask(MERKLE_CYCLE, [](size_t i) { return i < HASH_LEN && i % 8 < 7; }));
        push_ark(a.periodic);
        for (size_t i = 0; i < HRW; i++) a.assertions.push_back(single(PREV_ROOT + i, 0, canon(pub[i])));
        for (size_t i = 0; i < HRW; i++) a.assertions.push_back(single(PREV_ROOT + i, n - 1, canon(pub[7 + i])));
        break;
    }
    case 2: {   // PreMerkleAir
        need(npub == 29 && n == 16, "merkle-init AIR: 29 public inputs, trace length 16");
        a.width = 58;
        a.degrees.assign(4 * HSW, deg(3));
        push_ark(a.periodic);
        fe s[14], r[14], delta = canon(pub[28]);
        for (int i = 0; i < 14; i++) { s[i] = canon(pub[i]); r[i] = canon(pub[14 + i]); }
        for (size_t i = 0; i < APW + 2; i++) a.assertions.push_back(single(SENDER_INITIAL + i, 0, s[i]));
        for (size_t i = 0; i < APW; i++) a.assertions.push_back(single(SENDER_UPDATED + i, 0, s[i]));
        a.assertions.push_back(single(SENDER_UPDATED + APW, 0, sub(s[APW], delta)));
        a.assertions.push_back(single(SENDER_UPDATED + APW + 1, 0, add(s[APW + 1], ONE)));
        for (size_t i = 0; i < APW + 2; i++) a.assertions.push_back(single(RECEIVER_INITIAL + i, 0, r[i]));
        for (size_t i = 0; i < APW; i++) a.assertions.push_back(single(RECEIVER_UPDATED + i, 0, r[i]));
        a.assertions.push_back(single(RECEIVER_UPDATED + APW, 0, add(r[APW], delta)));
        a.assertions.push_back(single(RECEIVER_UPDATED + APW + 1, 0, r[APW + 1]));
        break;
    }
    case 3: {   // SchnorrAir
        need(npub % SIG_PUB_WORDS == 0 && npub > 0 && n == (npub / SIG_PUB_WORDS) * SIG_CYCLE, "schnorr AIR: 38 public words and 512 rows per signature");
        const size_t nsig = npub / SIG_PUB_WORDS;
        a.width = SCHNORR_WIDTH;
        a.degrees = schnorr_degrees(nsig, SIG_CYCLE);
        auto preds = schnorr_predicates();
        for (int i = 0; i < 7; i++) a.periodic.push_back(mask(SIG_CYCLE, preds[i]));
        for (size_t j = 0; j < APW; j++) {   // the signer's public key, constant over each signature: period = whole trace
            PeriodicColumn c; c.values.resize(n);
            for (size_t m = 0; m < nsig; m++) for (size_t i = 0; i < SIG_CYCLE; i++) c.values[m * SIG_CYCLE + i] = canon(pub[m * SIG_PUB_WORDS + j]);
            a.periodic.push_back(c);
        }
        a.periodic.push_back(mask(SIG_CYCLE, preds[7]));
        for (size_t j = 0; j < HRW; j++) {   // message chunks injected at steps 7, 15, 23, 31 of each signature
            PeriodicColumn c; c.values.assign(n, ZERO);
            for (size_t m = 0; m < nsig; m++) for (size_t i = 0; i < NUM_HASH_ITER - 1; i++) c.values[m * SIG_CYCLE + i * 8 + 7] = canon(pub[m * SIG_PUB_WORDS + i * HRW + j]);
            a.periodic.push_back(c);
        }
        push_ark(a.periodic);
        for (size_t i = 0; i < PPW; i++) a.assertions.push_back(periodic(i, 0, SIG_CYCLE, i == PCW ? ONE : ZERO));
        a.assertions.push_back(periodic(PPW, 0, SIG_CYCLE, ZERO));
        for (size_t i = 0; i < PPW; i++) a.assertions.push_back(periodic(i + PPW + 1, 0, SIG_CYCLE, i == PCW ? ONE : ZERO));
        for (size_t i = 0; i < 5; i++) a.assertions.push_back(periodic(i + 2 * PPW + 1, 0, SIG_CYCLE, ZERO));
        std::vector<std::vector<fe>> rx(PCW, std::vector<fe>(nsig));
        for (size_t l = 0; l < PCW; l++) for (size_t m = 0; m < nsig; m++) rx[l][m] = canon(pub[m * SIG_PUB_WORDS + 28 + l]);
        for (size_t l = 0; l < PCW; l++) a.assertions.push_back(sequence(2 * PPW + 6 + l, 0, SIG_CYCLE, rx[l]));
        for (size_t i = 0; i < HRW; i++) a.assertions.push_back(periodic(i + 2 * PPW + PCW + 6, 0, SIG_CYCLE, ZERO));
        for (size_t l = 0; l < PCW; l++) a.assertions.push_back(sequence(l, SCALAR_MUL_LENGTH + 1, SIG_CYCLE, rx[l]));
        break;
    }
    case 4: {   // RangeProofAir
        need(npub == 1, "range AIR: 1 public input");
        a.width = 2;
        a.degrees = {deg(2), deg(1)};
        a.assertions = {single(1, 0, ZERO), single(1, n - 1, canon(pub[0]))};
        break;
    }
    case 5: {   // RescueAir (benches/rescue.rs)
        need(npub == 14, "rescue AIR: 14 public inputs");
        a.width = 14;
        a.degrees.assign(HSW, deg(3, {8}));
        a.periodic.push_back(mask(8, [](size_t i) { return i < 7; }));
        push_ark(a.periodic);
        for (size_t i = 0; i < HRW; i++) a.assertions.push_back(single(i, 0, canon(pub[i])));
        for (size_t i = 0; i < HRW; i++) a.assertions.push_back(single(i, n - 1, canon(pub[7 + i])));
        break;
    }
    default: throw std::invalid_argument("unknown AIR id");
    }
    std::stable_sort(a.assertions.begin(), a.assertions.end(), [](const Assertion &x, const Assertion &y) {
        if (x.stride != y.stride) return x.stride < y.stride;
        if (x.first_step != y.first_step) return x.first_step < y.first_step;
        return x.column < y.column;
    });
    return a;
}

TransitionGroups transition_groups(const AirDesc &air) {
    TransitionGroups g;
    const size_t n = air.trace_len, comp_degree = n * air.ce_blowup() - 1, target = comp_degree + (n - 1);
    std::vector<size_t> degs;
    for (const Degree &d : air.degrees) {
        size_t ed = air.evaluation_degree(d);
        size_t gi = std::find(degs.begin(), degs.end(), ed) - degs.begin();
        if (gi == degs.size()) { degs.push_back(ed); g.adj.push_back(target - ed); }
        g.group_of.push_back((uint8_t)gi);
    }
    return g;
}
BoundaryGroups boundary_groups(const AirDesc &air) {
    BoundaryGroups b;
    const size_t n = air.trace_len, comp_degree = n * air.ce_blowup() - 1;
    const fe g = root_of_unity(ilog2_host(n));
    for (size_t i = 0; i < air.assertions.size(); i++) {
        const Assertion &s = air.assertions[i];
        if (i == 0 || s.stride != air.assertions[i - 1].stride || s.first_step != air.assertions[i - 1].first_step) {
            BoundaryGroup bg;
            bg.stride = s.stride; bg.first_step = s.first_step;
            bg.num_steps = s.stride == 0 ? 1 : n / s.stride;
            bg.offset = f63::pow(g, (uint64_t)bg.num_steps * s.first_step);
            bg.adj = comp_degree + bg.num_steps - (n - 1);
            b.groups.push_back(bg);
        }
        b.group_of.push_back((uint32_t)b.groups.size() - 1);
    }
    return b;
}

}  // namespace csg
