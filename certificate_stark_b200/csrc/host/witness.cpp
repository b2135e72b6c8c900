// Host-side witness generation: the build_trace() of each prover in the reference, plus a seeded stand-in for the
// OsRng-driven example generators (the reference's inputs are non-deterministic).
//   TransactionProver::build_trace   src/prover.rs:37-98, src/trace.rs:28-142
//   MerkleProver::build_trace        src/merkle/update/prover.rs:37-80, src/merkle/update/trace.rs
//   PreMerkleProver::build_trace     src/merkle/init/prover.rs:35-53, src/merkle/init/trace.rs
//   SchnorrProver::build_trace       src/schnorr/prover.rs:52-80, src/schnorr/trace.rs
//   RangeProver::build_trace         src/range/prover.rs:36-56
//   RescueProver::build_trace        benches/rescue.rs:279-321
//   TransactionMetadata::build_random src/lib.rs:235-464 ; SchnorrExample::new src/schnorr/mod.rs:79-141
// Traces are produced column-major in canonical form, the layout TraceTable stores and csg_prove() consumes.
// This is host code (the reference builds witnesses on the CPU too); moving it to the GPU is SURVEY.md 8(f).1.
#include <algorithm>
#include <array>
#include <cstring>
#include <vector>

#include <map>
#include <stdexcept>

#include "../../../include/csg.h"
#include "../ecc.cuh"
#include "../rescue.cuh"
#include "batch_plan.hpp"

using f63::fe;

namespace {

enum {
    HSW = 14, HRW = 7, APW = 12, PPW = 18, PCW = 6,
    SENDER_INITIAL_POS = 0, SENDER_BIT_POS = 14, SENDER_UPDATED_POS = 15, RECEIVER_INITIAL_POS = 29, RECEIVER_BIT_POS = 43,
    RECEIVER_UPDATED_POS = 44, PREV_TREE_ROOT_POS = 58, MERKLE_WIDTH = 65, MERKLE_CYCLE = 512,
    SENDER_KEY_POINT_POS = 65, DELTA_COPY_POS = 89, NONCE_COPY_POS = 91, TX_WIDTH = 94, TX_CYCLE = 1024,
    SCHNORR_WIDTH = 56, SIG_CYCLE = 512, SCALAR_MUL_LENGTH = 510, NUM_HASH_ITER = 5, RANGE_LOG = 64,
};

struct SplitMix64 {
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
};

typedef std::array<fe, 7> Hash7;
typedef std::array<fe, 14> Account;   // pk.x[6] pk.y[6] balance nonce (Montgomery)
typedef std::array<fe, 28> Message;   // sender pk, receiver pk, delta, nonce, 0, 0   (src/lib.rs:467-481)
struct U256 { uint64_t w[4]; };
struct Signature { std::array<fe, 6> rx; U256 s; };

inline bool bit(const U256 &v, unsigned i) { return (v.w[i >> 6] >> (i & 63)) & 1; }

// ---- curve helpers (host) ----
ecc::point identity() { ecc::point p{}; p.y.c[0] = f63::ONE; return p; }
ecc::point scalar_mul_generator(const U256 &k) {
    const ecc::fp6 gx = ecc::load6(CSG_GENERATOR_M), gy = ecc::load6(CSG_GENERATOR_M + 6);
    ecc::point acc = identity();
    for (int i = 254; i >= 0; i--) { acc = ecc::double_point(acc); if (bit(k, i)) acc = ecc::add_mixed(acc, gx, gy); }
    return acc;
}
void to_affine(const ecc::point &p, fe out[12]) {
    ecc::fp6 zi = ecc::inv(p.z), x = ecc::mul(p.x, zi), y = ecc::mul(p.y, zi);
    for (int i = 0; i < 6; i++) { out[i] = x.c[i]; out[6 + i] = y.c[i]; }
}

// h = merge(merge(merge(merge(digest(R.x), m[0..7]), m[7..14]), m[14..21]), m[21..28])   (src/schnorr/mod.rs:247-288)
Hash7 hash_message(const fe rx[6], const Message &m) {
    Hash7 h;
    rescue::digest(rx, 6, h.data());
    for (int k = 0; k < 4; k++) { Hash7 t; rescue::merge(h.data(), m.data() + 7 * k, t.data()); h = t; }
    return h;
}
U256 hash_to_scalar_bits(const Hash7 &h) { U256 v; for (int i = 0; i < 4; i++) v.w[i] = f63::from_mont(h[i]); return v; }  // schnorr/trace.rs:136-139

// r - k*h as plain integers; returns false if negative.  The circuit's double-and-add never reduces modulo the group
// order, so s*G + h*(k*G) = r*G holds for this s whatever the order is (the reference's Scalar type is absent here).
bool sub_small_multiple(const U256 &r, const U256 &h, unsigned k, U256 &out) {
    unsigned __int128 carry = 0; uint64_t kh[5];
    for (int i = 0; i < 4; i++) { unsigned __int128 t = (unsigned __int128)h.w[i] * k + carry; kh[i] = (uint64_t)t; carry = t >> 64; }
    kh[4] = (uint64_t)carry;
    if (kh[4]) return false;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) { unsigned __int128 t = (unsigned __int128)r.w[i] - kh[i] - borrow; out.w[i] = (uint64_t)t; borrow = (t >> 64) & 1; }
    return borrow == 0;
}
// Schnorr signature with secret key sk in {1,2,3} (public key sk*G)   (src/schnorr/mod.rs:197-216)
Signature sign(const Message &msg, unsigned sk, SplitMix64 &rng) {
    for (;;) {
        U256 r = {{rng.next(), rng.next(), rng.next(), rng.next()}};
        r.w[3] = (r.w[3] & 0x7fffffffffffffffULL) | 0x6000000000000000ULL;  // 2^254 + 2^253 <= r < 2^255
        fe aff[12];
        to_affine(scalar_mul_generator(r), aff);
        Signature sig;
        for (int i = 0; i < 6; i++) sig.rx[i] = aff[i];
        U256 h = hash_to_scalar_bits(hash_message(aff, msg));
        if (sub_small_multiple(r, h, sk, sig.s)) return sig;
    }
}

// ---- Rescue Merkle tree of accounts (winterfell MerkleTree<Rescue63>: node i = merge(node 2i, node 2i+1)) ----
struct AccountTree {
    unsigned depth; std::vector<Hash7> nodes;
    explicit AccountTree(unsigned d) : depth(d), nodes((size_t)2 << d) {
        Hash7 z{}; z.fill(0);
        size_t nl = (size_t)1 << d;
        for (size_t i = nl; i < 2 * nl; i++) nodes[i] = z;
        for (unsigned l = d; l-- > 0;) {  // all nodes of a level are equal in the empty tree
            Hash7 v; rescue::merge(nodes[(size_t)2 << l].data(), nodes[((size_t)2 << l) + 1].data(), v.data());
            for (size_t i = (size_t)1 << l; i < (size_t)2 << l; i++) nodes[i] = v;
        }
    }
    void update_leaf(size_t idx, const Hash7 &leaf) {
        size_t i = ((size_t)1 << depth) + idx; nodes[i] = leaf;
        for (i >>= 1; i >= 1; i >>= 1) rescue::merge(nodes[2 * i].data(), nodes[2 * i + 1].data(), nodes[i].data());
    }
    // bulk form for the initial accounts: leaves are written first, then every touched interior node is recomputed once, level
    // by level, on all host threads (thousands of independent Rescue permutations instead of 15 dependent ones per account)
    std::vector<size_t> dirty;
    void set_leaf(size_t idx, const Hash7 &leaf) { size_t i = ((size_t)1 << depth) + idx; nodes[i] = leaf; dirty.push_back(i); }
    void rebuild() {
        std::vector<size_t> cur;
        cur.swap(dirty);
        for (unsigned l = depth; l-- > 0;) {
            for (auto &i : cur) i >>= 1;
            std::sort(cur.begin(), cur.end());
            cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
#pragma omp parallel for schedule(static)
            for (long k = 0; k < (long)cur.size(); k++) rescue::merge(nodes[2 * cur[k]].data(), nodes[2 * cur[k] + 1].data(), nodes[cur[k]].data());
        }
    }
    std::vector<Hash7> prove(size_t idx) const {  // [leaf, sibling leaf, sibling nodes ... up to below the root]
        std::vector<Hash7> p; size_t i = ((size_t)1 << depth) + idx;
        p.push_back(nodes[i]);
        for (; i > 1; i >>= 1) p.push_back(nodes[i ^ 1]);
        return p;
    }
    const Hash7 &root() const { return nodes[1]; }
};
Hash7 account_leaf(const Account &a) { Hash7 h; rescue::merge(a.data(), a.data() + 7, h.data()); return h; }

}  // namespace

struct csg_tx_batch {
    unsigned tree_depth;
    std::vector<Hash7> initial_roots; Hash7 final_root;
    std::vector<Account> s_old, r_old;
    std::vector<size_t> s_idx, r_idx;
    std::vector<std::vector<Hash7>> s_paths, r_paths;
    std::vector<fe> deltas;
    std::vector<Signature> sigs;
    std::vector<Message> msgs;
};
struct csg_sig_batch { std::vector<Message> msgs; std::vector<Signature> sigs; };

namespace {

Message build_tx_message(const Account &s, const Account &r, fe delta, fe nonce) {
    Message m; m.fill(0);
    for (int i = 0; i < APW; i++) { m[i] = s[i]; m[APW + i] = r[i]; }
    m[2 * APW] = delta; m[2 * APW + 1] = nonce;
    return m;
}

// ---- per-row state updates ----
void merkle_auth_step(size_t pos, size_t index, const std::vector<Hash7> &branch, fe *st /* 29 wide */) {  // update/trace.rs:97-136
    size_t cyc = pos / 8, cp = pos % 8;
    if (cp < 7) { rescue::apply_round(st, pos); rescue::apply_round(st + HSW + 1, pos); }
    else {
        const Hash7 &node = branch[cyc + 1];
        bool b = (index >> cyc) & 1;
        for (int i = 0; i < HRW; i++) {
            if (!b) { st[HRW + i] = node[i]; st[HSW + 1 + HRW + i] = node[i]; }
            else { st[HRW + i] = st[i]; st[HSW + 1 + HRW + i] = st[HSW + 1 + i]; st[i] = node[i]; st[HSW + 1 + i] = node[i]; }
        }
        st[HSW] = b ? f63::ONE : 0;
    }
}
void merkle_update_init(const Hash7 &root, const Account &s, const Account &r, fe delta, fe *st) {  // update/trace.rs:19-48
    for (int i = 0; i < 14; i++) { st[SENDER_INITIAL_POS + i] = s[i]; st[SENDER_UPDATED_POS + i] = s[i]; st[RECEIVER_INITIAL_POS + i] = r[i]; st[RECEIVER_UPDATED_POS + i] = r[i]; }
    st[SENDER_BIT_POS] = 0; st[RECEIVER_BIT_POS] = 0;
    st[SENDER_UPDATED_POS + APW] = f63::sub(st[SENDER_UPDATED_POS + APW], delta);
    st[SENDER_UPDATED_POS + APW + 1] = f63::add(st[SENDER_UPDATED_POS + APW + 1], f63::ONE);
    st[RECEIVER_UPDATED_POS + APW] = f63::add(st[RECEIVER_UPDATED_POS + APW], delta);
    for (int i = 0; i < HRW; i++) st[PREV_TREE_ROOT_POS + i] = root[i];
}
void merkle_update_step(size_t step, unsigned depth, size_t si, size_t ri, const std::vector<Hash7> &sb, const std::vector<Hash7> &rb, fe *st) {  // update/trace.rs:53-95
    size_t hash_len = 8 * depth + 7;
    if (step < hash_len) { merkle_auth_step(step, si, sb, st + SENDER_INITIAL_POS); merkle_auth_step(step, ri, rb, st + RECEIVER_INITIAL_POS); }
    if (step == hash_len - 1) for (int i = 0; i < HRW; i++) st[PREV_TREE_ROOT_POS + i] = st[RECEIVER_UPDATED_POS + i];
}
void store_point(const ecc::point &p, fe *st) { for (int i = 0; i < 6; i++) { st[i] = p.x.c[i]; st[6 + i] = p.y.c[i]; st[12 + i] = p.z.c[i]; } }
ecc::point load_point(const fe *st) { return {ecc::load6(st), ecc::load6(st + 6), ecc::load6(st + 12)}; }
void schnorr_init(const Signature &sig, fe *st) {  // schnorr/trace.rs:18-30
    for (int i = 0; i < SCHNORR_WIDTH; i++) st[i] = 0;
    st[PCW] = f63::ONE; st[PPW + PCW + 1] = f63::ONE;
    for (int i = 0; i < 6; i++) st[2 * PPW + 6 + i] = sig.rx[i];
}
void schnorr_step(size_t step, const Message &msg, const U256 &s_bits, const U256 &h_bits, fe *st) {  // schnorr/trace.rs:35-122
    const size_t H = 2 * PPW + 6;
    bool rescue_flag = step < 8 * NUM_HASH_ITER;
    if (rescue_flag && step % 8 < 7) rescue::apply_round(st + H, step);
    else if (rescue_flag && step < (NUM_HASH_ITER - 1) * 8) for (int i = 0; i < HRW; i++) st[H + HRW + i] = msg[HRW * (step / 8) + i];
    else if (rescue_flag) for (int i = 0; i < HRW; i++) st[H + HRW + i] = 0;
    if (step < SCALAR_MUL_LENGTH) {
        size_t real = step / 2, chunk = real < 63 ? 0 : (real - 63) / 64 + 1;
        st[PPW] = bit(s_bits, 254 - real) ? f63::ONE : 0;
        st[2 * PPW + 1] = bit(h_bits, 254 - real) ? f63::ONE : 0;
        if (step % 2 == 0) {
            store_point(ecc::double_point(load_point(st)), st);
            store_point(ecc::double_point(load_point(st + PPW + 1)), st + PPW + 1);
            fe *acc = st + 2 * PPW + 1 + (4 - chunk);
            *acc = f63::add(f63::dbl(*acc), st[2 * PPW + 1]);
        } else {
            if (st[PPW] == f63::ONE) store_point(ecc::add_mixed(load_point(st), ecc::load6(CSG_GENERATOR_M), ecc::load6(CSG_GENERATOR_M + 6)), st);
            if (st[2 * PPW + 1] == f63::ONE) store_point(ecc::add_mixed(load_point(st + PPW + 1), ecc::load6(msg.data()), ecc::load6(msg.data() + 6)), st + PPW + 1);
        }
    } else if (step == SCALAR_MUL_LENGTH) {
        st[PPW] = f63::ONE;
        ecc::point sum = ecc::add_full(load_point(st), load_point(st + PPW + 1));
        store_point(sum, st);
        ecc::fp6 x = ecc::mul(sum.x, ecc::inv(sum.z));
        for (int i = 0; i < 6; i++) st[i] = x.c[i];
    }
}
U256 le_bits(fe v) { U256 r = {{f63::from_mont(v), 0, 0, 0}}; return r; }
// TraceTable::fill: row 0 <- init, row i+1 <- update(i, copy of row i); written column-major canonical
template <class Init, class Update>
void fill_fragment(uint64_t *trace, size_t total_len, size_t width, size_t row0, size_t len, Init init, Update update) {
    std::vector<fe> st(width, 0);
    init(st.data());
    for (size_t c = 0; c < width; c++) trace[c * total_len + row0] = f63::from_mont(st[c]);
    for (size_t i = 0; i + 1 < len; i++) {
        update(i, st.data());
        for (size_t c = 0; c < width; c++) trace[c * total_len + row0 + i + 1] = f63::from_mont(st[c]);
    }
}
// ---- the seeded draws of TransactionMetadata::build_random (src/lib.rs:235-464), with no hashing: shared by the host builder
// below and by the plan of the device-side builder
struct BatchDraws {
    std::vector<size_t> created; std::vector<Account> created_values;   // initial accounts in draw order (a slot may repeat: the last one stays)
    std::vector<size_t> s_idx, r_idx;
    std::vector<Account> s_old, r_old, s_new, r_new;                     // the two leaves before and after each transfer
    std::vector<fe> deltas;
    std::vector<unsigned> s_sk;
    std::vector<uint64_t> sig_seeds;
};
BatchDraws draw_batch(uint64_t seed, size_t num_tx, unsigned tree_depth) {
    SplitMix64 rng{seed};
    BatchDraws D;
    const size_t tree_size = (size_t)1 << tree_depth;
    fe pk[4][12];
    for (unsigned k = 1; k <= 3; k++) { U256 s = {{k, 0, 0, 0}}; to_affine(scalar_mul_generator(s), pk[k]); }
    std::vector<unsigned> skeys(tree_size, 0);
    std::vector<Account> values(tree_size); for (auto &v : values) v.fill(0);
    auto new_account = [&](size_t idx) {
        unsigned sk = 1 + rng.next() % 3; skeys[idx] = sk;
        Account a;
        for (int i = 0; i < 12; i++) a[i] = pk[sk][i];
        a[12] = f63::to_mont(rng.next() % f63::P); a[13] = f63::to_mont(rng.next() % f63::P);
        values[idx] = a; D.created.push_back(idx); D.created_values.push_back(a);
    };
    D.s_idx.resize(num_tx); D.r_idx.resize(num_tx);
    for (size_t t = 0; t < num_tx; t++) { D.s_idx[t] = rng.next() % tree_size; new_account(D.s_idx[t]); }
    for (size_t t = 0; t < num_tx; t++) {
        size_t r = rng.next() % tree_size;
        while (r == D.s_idx[t]) r = rng.next() % tree_size;
        D.r_idx[t] = r;
        if (!skeys[r]) new_account(r);
    }
    for (size_t t = 0; t < num_tx; t++) {
        const size_t si = D.s_idx[t], ri = D.r_idx[t];
        uint64_t sb = f63::from_mont(values[si][12]), rb = f63::from_mont(values[ri][12]);
        uint64_t bound = std::min(sb, UINT64_MAX - rb);
        fe delta = f63::to_mont((rng.next() % (bound ? bound : 1)) % f63::P);
        D.s_sk.push_back(skeys[si]); D.s_old.push_back(values[si]); D.r_old.push_back(values[ri]); D.deltas.push_back(delta);
        values[si][12] = f63::sub(values[si][12], delta); values[si][13] = f63::add(values[si][13], f63::ONE);
        values[ri][12] = f63::add(values[ri][12], delta);
        D.s_new.push_back(values[si]); D.r_new.push_back(values[ri]);
    }
    D.sig_seeds.resize(num_tx);
    for (auto &s : D.sig_seeds) s = rng.next();
    return D;
}
}  // namespace

// the tree's history as merge records (batch_plan.hpp): the AccountTree above, with record ids in place of hashes
csg::BatchPlan csg::plan_tx_batch(uint64_t seed, size_t num_tx, unsigned tree_depth) {
    if (!num_tx || tree_depth < 1 || tree_depth > 15 || ((tree_depth + 1) & tree_depth)) throw std::invalid_argument("bad transaction batch parameters");
    const BatchDraws D = draw_batch(seed, num_tx, tree_depth);
    BatchPlan P;
    P.depth = tree_depth; P.ntx = num_tx;
    const size_t nev = 2 * num_tx;
    // time 0: the last account drawn for each slot, then the set of touched nodes level by level
    std::vector<std::map<size_t, int32_t>> last(tree_depth + 1);   // node index -> latest record id, per level
    std::map<size_t, size_t> base_leaf;                              // slot -> index into created (last wins)
    for (size_t k = 0; k < D.created.size(); k++) base_leaf[D.created[k]] = k;
    P.nbase.assign(tree_depth + 1, 0);
    P.nbase[0] = (uint32_t)base_leaf.size();
    std::vector<std::vector<size_t>> base_nodes(tree_depth + 1);
    for (auto &kv : base_leaf) base_nodes[0].push_back(kv.first);
    for (unsigned l = 1; l <= tree_depth; l++) {
        for (size_t i : base_nodes[l - 1]) if (base_nodes[l].empty() || base_nodes[l].back() != (i >> 1)) base_nodes[l].push_back(i >> 1);   // sorted input: duplicates are adjacent
        P.nbase[l] = (uint32_t)base_nodes[l].size();
    }
    P.level_off.assign(tree_depth + 2, 0);
    for (unsigned l = 0; l <= tree_depth; l++) P.level_off[l + 1] = P.level_off[l] + P.nbase[l] + (uint32_t)nev;
    const size_t total = P.level_off[tree_depth + 1];
    P.left.assign(total, 0); P.right.assign(total, 0);
    P.accounts.resize((size_t)14 * (P.nbase[0] + nev));
    {
        size_t k = 0;
        for (auto &kv : base_leaf) { memcpy(&P.accounts[14 * k], D.created_values[kv.second].data(), 14 * sizeof(fe)); last[0][kv.first] = (int32_t)(P.level_off[0] + k); k++; }
    }
    auto ref = [&](unsigned level, size_t node) -> int32_t { auto it = last[level].find(node); return it == last[level].end() ? -(int32_t)(level + 1) : it->second; };
    for (unsigned l = 1; l <= tree_depth; l++)
        for (size_t k = 0; k < base_nodes[l].size(); k++) {
            const size_t node = base_nodes[l][k];
            const int32_t id = (int32_t)(P.level_off[l] + k);
            P.left[id] = ref(l - 1, 2 * node); P.right[id] = ref(l - 1, 2 * node + 1);
            last[l][node] = id;
        }
    // the update events in time order
    P.tx_words.resize((size_t)BatchPlan::TX_WORDS * num_tx); P.tx_refs.resize((size_t)BatchPlan::TX_REFS * num_tx);
    auto path = [&](size_t leaf, int32_t *out) {   // [leaf version, sibling version at level 0, 1, ...] (winterfell MerkleTree::prove)
        out[0] = ref(0, leaf);
        for (unsigned l = 0; l < tree_depth; l++) out[1 + l] = ref(l, (leaf >> l) ^ 1);
        for (unsigned l = tree_depth + 1; l < 16; l++) out[l] = -1;
    };
    auto update = [&](size_t ev, size_t leaf, const Account &value) {
        memcpy(&P.accounts[14 * (P.nbase[0] + ev)], value.data(), 14 * sizeof(fe));
        for (unsigned l = 0; l <= tree_depth; l++) {
            const size_t node = leaf >> l;
            const int32_t id = (int32_t)(P.level_off[l] + P.nbase[l] + ev);
            if (l) {   // the child on the path was just written; the other one is whatever version is current
                const size_t child = leaf >> (l - 1);
                const int32_t own = (int32_t)(P.level_off[l - 1] + P.nbase[l - 1] + ev), other = ref(l - 1, child ^ 1);
                P.left[id] = (child & 1) ? other : own; P.right[id] = (child & 1) ? own : other;
            }
            last[l][node] = id;
        }
    };
    for (size_t t = 0; t < num_tx; t++) {
        uint64_t *w = &P.tx_words[(size_t)BatchPlan::TX_WORDS * t];
        int32_t *r = &P.tx_refs[(size_t)BatchPlan::TX_REFS * t];
        memcpy(w, D.s_old[t].data(), 14 * sizeof(fe)); memcpy(w + 14, D.r_old[t].data(), 14 * sizeof(fe));
        w[28] = D.deltas[t]; w[29] = D.s_idx[t]; w[30] = D.r_idx[t]; w[31] = D.s_sk[t]; w[32] = D.sig_seeds[t];
        r[32] = ref(tree_depth, 0);
        path(D.s_idx[t], r);
        update(2 * t, D.s_idx[t], D.s_new[t]);
        update(2 * t + 1, D.r_idx[t], D.r_new[t]);
        path(D.r_idx[t], r + 16);
    }
    P.final_root = ref(tree_depth, 0);
    return P;
}

extern "C" {

int csg_build_trace_rescue(const uint64_t seed[7], size_t chain_length, uint64_t *trace, uint64_t pub[14]) {
    if (!chain_length || (chain_length & (chain_length - 1))) return CSG_ERR_ARG;
    size_t n = chain_length * 8;
    fill_fragment(trace, n, 14, 0, n,
        [&](fe *st) { for (int i = 0; i < 7; i++) { st[i] = f63::to_mont(seed[i] % f63::P); st[7 + i] = 0; } },
        [&](size_t step, fe *st) { if (step % 8 < 7) rescue::apply_round(st, step); else for (int i = 7; i < 14; i++) st[i] = 0; });
    for (int i = 0; i < 7; i++) { pub[i] = trace[(size_t)i * n]; pub[7 + i] = trace[(size_t)i * n + n - 1]; }
    return CSG_OK;
}

int csg_build_trace_range(uint64_t number, uint64_t *trace, uint64_t pub[1]) {
    uint64_t v = number % f63::P;
    fill_fragment(trace, RANGE_LOG, 2, 0, RANGE_LOG, [&](fe *st) { st[0] = st[1] = 0; },
        [&](size_t step, fe *st) {  // range/prover.rs:74-84 called with range_log - 1
            if (step < RANGE_LOG - 1) { st[0] = ((v >> (RANGE_LOG - 2 - step)) & 1) ? f63::ONE : 0; st[1] = f63::add(f63::dbl(st[1]), st[0]); }
        });
    pub[0] = trace[RANGE_LOG + RANGE_LOG - 1];
    return CSG_OK;
}

int csg_build_trace_merkle_init(const uint64_t s_inputs[14], const uint64_t r_inputs[14], uint64_t delta, uint64_t *trace, uint64_t pub[29]) {
    const size_t n = 16, w = 58;
    fe d = f63::to_mont(delta % f63::P);
    fill_fragment(trace, n, w, 0, n,
        [&](fe *st) {  // init/trace.rs:19-47 (including its habit of writing the sender's coins/nonce into the UPDATED slots first)
            fe s[14], r[14];
            for (int i = 0; i < 14; i++) { s[i] = f63::to_mont(s_inputs[i] % f63::P); r[i] = f63::to_mont(r_inputs[i] % f63::P); }
            for (int i = 0; i < APW; i++) { st[SENDER_INITIAL_POS + i] = s[i]; st[SENDER_UPDATED_POS + i] = s[i]; st[RECEIVER_INITIAL_POS + i] = r[i]; st[RECEIVER_UPDATED_POS + i] = r[i]; }
            st[SENDER_UPDATED_POS + APW] = f63::sub(s[APW], d); st[SENDER_UPDATED_POS + APW + 1] = f63::add(s[APW + 1], f63::ONE);
            st[RECEIVER_INITIAL_POS + APW] = r[APW]; st[RECEIVER_INITIAL_POS + APW + 1] = r[APW + 1];
            st[RECEIVER_UPDATED_POS + APW] = f63::add(r[APW], d); st[RECEIVER_UPDATED_POS + APW + 1] = r[APW + 1];
        },
        [&](size_t step, fe *st) { rescue::apply_round(st + SENDER_INITIAL_POS, step); rescue::apply_round(st + SENDER_UPDATED_POS, step);
                                   rescue::apply_round(st + RECEIVER_INITIAL_POS, step); rescue::apply_round(st + RECEIVER_UPDATED_POS, step); });
    // PreMerkleProver::get_pub_inputs reads them back from row 0 (init/prover.rs:60-100)
    for (int i = 0; i < 14; i++) { pub[i] = trace[(size_t)(SENDER_INITIAL_POS + i) * n]; pub[14 + i] = trace[(size_t)(RECEIVER_INITIAL_POS + i) * n]; }
    fe ru = f63::to_mont(trace[(size_t)(RECEIVER_UPDATED_POS + APW) * n]), ri = f63::to_mont(trace[(size_t)(RECEIVER_INITIAL_POS + APW) * n]);
    pub[28] = f63::from_mont(f63::sub(ru, ri));
    return CSG_OK;
}

csg_tx_batch *csg_tx_batch_new(uint64_t seed, size_t num_tx, unsigned tree_depth) {
    if (!num_tx || tree_depth < 1 || tree_depth > 15 || ((tree_depth + 1) & tree_depth)) return nullptr;  // depth+1 must be a power of two (src/lib.rs:106-109)
    const BatchDraws D = draw_batch(seed, num_tx, tree_depth);
    auto *B = new csg_tx_batch;
    B->tree_depth = tree_depth;
    B->s_idx = D.s_idx; B->r_idx = D.r_idx;
    AccountTree tree(tree_depth);
    {   // a slot drawn twice keeps its last account, as with one update per draw
        std::vector<Hash7> leaves(D.created.size());
#pragma omp parallel for schedule(static)
        for (long k = 0; k < (long)D.created.size(); k++) leaves[k] = account_leaf(D.created_values[k]);
        for (size_t k = 0; k < D.created.size(); k++) tree.set_leaf(D.created[k], leaves[k]);
        tree.rebuild();
    }
    for (size_t t = 0; t < num_tx; t++) {
        const size_t si = D.s_idx[t], ri = D.r_idx[t];
        B->initial_roots.push_back(tree.root());
        B->s_old.push_back(D.s_old[t]); B->r_old.push_back(D.r_old[t]); B->deltas.push_back(D.deltas[t]);
        B->s_paths.push_back(tree.prove(si));
        tree.update_leaf(si, account_leaf(D.s_new[t])); tree.update_leaf(ri, account_leaf(D.r_new[t]));
        B->r_paths.push_back(tree.prove(ri));
    }
    B->final_root = tree.root();
    B->msgs.resize(num_tx); B->sigs.resize(num_tx);
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < num_tx; t++) {
        SplitMix64 r2{D.sig_seeds[t]};
        B->msgs[t] = build_tx_message(B->s_old[t], B->r_old[t], B->deltas[t], B->s_old[t][13]);
        B->sigs[t] = sign(B->msgs[t], D.s_sk[t], r2);
    }
    return B;
}
// Test helper (no GPU needed): executes the device builder's plan (batch_plan.hpp) on the host, step for step what batch_gen.cu's
// kernels do -- leaf hashes, one merge per node version level by level, path gathers, signatures -- and writes the packed records.
// tests/test_host.py compares them with csg_tx_batch_pack of the sequential host builder: that pins the PLAN (which versions exist,
// which children they merge, which versions the paths read) in the CPU suite; the kernels themselves are pinned on the GPU.
int csg_debug_tx_batch_plan_records(uint64_t seed, size_t num_tx, unsigned tree_depth, uint64_t *out /* 278 words per transfer */, uint64_t pub[14]) {
    csg::BatchPlan P;
    try { P = csg::plan_tx_batch(seed, num_tx, tree_depth); } catch (const std::exception &) { return CSG_ERR_ARG; }
    const size_t total = P.level_off[P.depth + 1];
    std::vector<Hash7> hashes(total), defaults(P.depth + 1);
    defaults[0].fill(0);
    for (unsigned l = 0; l < P.depth; l++) rescue::merge(defaults[l].data(), defaults[l].data(), defaults[l + 1].data());
    auto version = [&](int32_t ref) -> const Hash7 & { return ref >= 0 ? hashes[ref] : defaults[-ref - 1]; };
#pragma omp parallel for schedule(static)
    for (long g = 0; g < (long)P.level_off[1]; g++) rescue::merge(&P.accounts[14 * g], &P.accounts[14 * g + 7], hashes[g].data());
    for (unsigned l = 1; l <= P.depth; l++) {
#pragma omp parallel for schedule(static)
        for (long id = P.level_off[l]; id < (long)P.level_off[l + 1]; id++) rescue::merge(version(P.left[id]).data(), version(P.right[id]).data(), hashes[id].data());
    }
    const size_t W = 278;
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < num_tx; t++) {
        const uint64_t *w = &P.tx_words[(size_t)csg::BatchPlan::TX_WORDS * t];
        const int32_t *rf = &P.tx_refs[(size_t)csg::BatchPlan::TX_REFS * t];
        uint64_t *r = out + t * W;
        memset(r, 0, W * sizeof(uint64_t));
        for (int i = 0; i < 28; i++) r[i] = w[i];
        r[28] = w[28]; r[36] = w[29]; r[37] = w[30];
        for (int i = 0; i < 7; i++) r[29 + i] = version(rf[32])[i];
        for (int k = 0; k < 32; k++) for (int i = 0; i < 7; i++) r[38 + 7 * k + i] = version(rf[k])[i];
        Account s_old, r_old;
        for (int i = 0; i < 14; i++) { s_old[i] = w[i]; r_old[i] = w[14 + i]; }
        const Message msg = build_tx_message(s_old, r_old, w[28], s_old[13]);
        SplitMix64 r2{w[32]};
        const Signature sig = sign(msg, (unsigned)w[31], r2);
        for (int i = 0; i < 6; i++) r[262 + i] = sig.rx[i];
        for (int i = 0; i < 4; i++) r[268 + i] = sig.s.w[i];
        U256 h = hash_to_scalar_bits(hash_message(sig.rx.data(), msg));
        for (int i = 0; i < 4; i++) r[272 + i] = h.w[i];
    }
    for (int i = 0; i < 7; i++) { pub[i] = f63::from_mont(version(P.tx_refs[32])[i]); pub[7 + i] = f63::from_mont(version(P.final_root)[i]); }
    return CSG_OK;
}
void csg_tx_batch_free(csg_tx_batch *b) { delete b; }
size_t csg_tx_batch_size(const csg_tx_batch *b) { return b->deltas.size(); }
void csg_tx_batch_roots(const csg_tx_batch *b, uint64_t initial_root[7], uint64_t final_root[7]) {
    for (int i = 0; i < 7; i++) { initial_root[i] = f63::from_mont(b->initial_roots[0][i]); final_root[i] = f63::from_mont(b->final_root[i]); }
}

// packed inputs of the device-side witness builder (csrc/witness.cuh): WIT_WORDS words per transaction
unsigned csg_tx_batch_depth(const csg_tx_batch *b) { return b->tree_depth; }
size_t csg_tx_batch_pack(const csg_tx_batch *B, uint64_t *out /* 278 words per transaction, or NULL for the size */) {
    const size_t W = 278, ntx = B->deltas.size();
    if (!out) return W * ntx;
#pragma omp parallel for schedule(static)
    for (size_t t = 0; t < ntx; t++) {
        uint64_t *r = out + t * W;
        memset(r, 0, W * sizeof(uint64_t));
        for (int i = 0; i < 14; i++) { r[i] = B->s_old[t][i]; r[14 + i] = B->r_old[t][i]; }
        r[28] = B->deltas[t];
        for (int i = 0; i < 7; i++) r[29 + i] = B->initial_roots[t][i];
        r[36] = B->s_idx[t]; r[37] = B->r_idx[t];
        for (size_t k = 0; k < B->s_paths[t].size() && k < 16; k++)
            for (int i = 0; i < 7; i++) { r[38 + 7 * k + i] = B->s_paths[t][k][i]; r[150 + 7 * k + i] = B->r_paths[t][k][i]; }
        for (int i = 0; i < 6; i++) r[262 + i] = B->sigs[t].rx[i];
        for (int i = 0; i < 4; i++) r[268 + i] = B->sigs[t].s.w[i];
        U256 h = hash_to_scalar_bits(hash_message(B->sigs[t].rx.data(), B->msgs[t]));
        for (int i = 0; i < 4; i++) r[272 + i] = h.w[i];
    }
    return W * ntx;
}

// the same record for a standalone signature: the message fills the slots the transfer's message is assembled from (witness.cuh)
size_t csg_sig_batch_pack(const csg_sig_batch *B, uint64_t *out /* 278 words per signature, or NULL for the size */) {
    const size_t W = 278, n = B->sigs.size();
    if (!out) return W * n;
#pragma omp parallel for schedule(static)
    for (size_t t = 0; t < n; t++) {
        uint64_t *r = out + t * W;
        const Message &m = B->msgs[t];
        memset(r, 0, W * sizeof(uint64_t));
        for (int i = 0; i < 12; i++) { r[i] = m[i]; r[14 + i] = m[12 + i]; }
        r[28] = m[24]; r[13] = m[25]; r[276] = m[26]; r[277] = m[27];
        for (int i = 0; i < 6; i++) r[262 + i] = B->sigs[t].rx[i];
        for (int i = 0; i < 4; i++) r[268 + i] = B->sigs[t].s.w[i];
        U256 h = hash_to_scalar_bits(hash_message(B->sigs[t].rx.data(), m));
        for (int i = 0; i < 4; i++) r[272 + i] = h.w[i];
    }
    return W * n;
}

int csg_build_trace_transaction(const csg_tx_batch *B, uint64_t *trace, uint64_t pub[14]) {
    size_t ntx = B->deltas.size(), n = ntx * TX_CYCLE;
    if (ntx & (ntx - 1)) return CSG_ERR_ARG;
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < ntx; t++) {
        const Account &s = B->s_old[t], &r = B->r_old[t];
        fe delta = B->deltas[t], sigma = f63::sub(s[12], delta);
        U256 dbits = le_bits(delta), sbits = le_bits(sigma), hbits = hash_to_scalar_bits(hash_message(B->sigs[t].rx.data(), B->msgs[t]));
        fill_fragment(trace, n, TX_WIDTH, t * TX_CYCLE, TX_CYCLE,
            [&](fe *st) {  // src/trace.rs:28-53
                merkle_update_init(B->initial_roots[t], s, r, delta, st);
                for (int i = 0; i < APW; i++) { st[SENDER_KEY_POINT_POS + i] = s[i]; st[SENDER_KEY_POINT_POS + APW + i] = r[i]; }
                st[DELTA_COPY_POS] = delta; st[DELTA_COPY_POS + 1] = sigma; st[NONCE_COPY_POS] = s[13];
            },
            [&](size_t step, fe *st) {  // src/trace.rs:59-142
                if (step < MERKLE_CYCLE - 1) merkle_update_step(step, B->tree_depth, B->s_idx[t], B->r_idx[t], B->s_paths[t], B->r_paths[t], st);
                else if (step == MERKLE_CYCLE - 1) { schnorr_init(B->sigs[t], st); st[SCHNORR_WIDTH] = st[SCHNORR_WIDTH + 1] = 0; st[NONCE_COPY_POS + 1] = st[NONCE_COPY_POS + 2] = 0; }
                else {
                    size_t ss = step - MERKLE_CYCLE;
                    schnorr_step(ss, B->msgs[t], B->sigs[t].s, hbits, st);
                    if (ss < RANGE_LOG) {  // range/prover.rs:74-84 with range_log = 64
                        fe *d = st + SCHNORR_WIDTH, *g = st + NONCE_COPY_POS + 1;
                        d[0] = bit(dbits, RANGE_LOG - 1 - ss) ? f63::ONE : 0; d[1] = f63::add(f63::dbl(d[1]), d[0]);
                        g[0] = bit(sbits, RANGE_LOG - 1 - ss) ? f63::ONE : 0; g[1] = f63::add(f63::dbl(g[1]), g[0]);
                    }
                }
            });
    }
    for (int i = 0; i < 7; i++) { pub[i] = trace[(size_t)(PREV_TREE_ROOT_POS + i) * n]; pub[7 + i] = trace[(size_t)(PREV_TREE_ROOT_POS + i) * n + n - 1]; }
    return CSG_OK;
}

int csg_build_trace_merkle_update(const csg_tx_batch *B, uint64_t *trace, uint64_t pub[14]) {
    size_t ntx = B->deltas.size(), n = ntx * MERKLE_CYCLE;
    if (ntx & (ntx - 1)) return CSG_ERR_ARG;
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < ntx; t++)
        fill_fragment(trace, n, MERKLE_WIDTH, t * MERKLE_CYCLE, MERKLE_CYCLE,
            [&](fe *st) { merkle_update_init(B->initial_roots[t], B->s_old[t], B->r_old[t], B->deltas[t], st); },
            [&](size_t step, fe *st) { merkle_update_step(step, B->tree_depth, B->s_idx[t], B->r_idx[t], B->s_paths[t], B->r_paths[t], st); });
    trace[(size_t)SENDER_BIT_POS * n + 1] = 1; trace[(size_t)RECEIVER_BIT_POS * n + 1] = 1;  // update/prover.rs:76-77
    for (int i = 0; i < 7; i++) { pub[i] = trace[(size_t)(PREV_TREE_ROOT_POS + i) * n]; pub[7 + i] = trace[(size_t)(PREV_TREE_ROOT_POS + i) * n + n - 1]; }
    return CSG_OK;
}

csg_sig_batch *csg_sig_batch_new(uint64_t seed, size_t num_sig) {
    if (!num_sig) return nullptr;
    SplitMix64 rng{seed};
    auto *B = new csg_sig_batch;
    fe pk[4][12];
    for (unsigned k = 1; k <= 3; k++) { U256 s = {{k, 0, 0, 0}}; to_affine(scalar_mul_generator(s), pk[k]); }
    std::vector<unsigned> sk(num_sig); std::vector<uint64_t> seeds(num_sig);
    B->msgs.resize(num_sig); B->sigs.resize(num_sig);
    for (size_t i = 0; i < num_sig; i++) {  // src/schnorr/mod.rs:86-101: message = pkey || 16 random elements
        sk[i] = 1 + rng.next() % 3;
        for (int j = 0; j < 12; j++) B->msgs[i][j] = pk[sk[i]][j];
        for (int j = 12; j < 28; j++) B->msgs[i][j] = f63::to_mont(rng.next() % f63::P);
        seeds[i] = rng.next();
    }
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < num_sig; i++) { SplitMix64 r2{seeds[i]}; B->sigs[i] = sign(B->msgs[i], sk[i], r2); }
    return B;
}
void csg_sig_batch_free(csg_sig_batch *b) { delete b; }
size_t csg_sig_batch_size(const csg_sig_batch *b) { return b->sigs.size(); }

int csg_build_trace_schnorr(const csg_sig_batch *B, uint64_t *trace, uint64_t *pub /* 38 per signature */) {
    size_t ns = B->sigs.size(), n = ns * SIG_CYCLE;
    if (ns & (ns - 1)) return CSG_ERR_ARG;
#pragma omp parallel for schedule(dynamic)
    for (size_t t = 0; t < ns; t++) {
        U256 hbits = hash_to_scalar_bits(hash_message(B->sigs[t].rx.data(), B->msgs[t]));
        fill_fragment(trace, n, SCHNORR_WIDTH, t * SIG_CYCLE, SIG_CYCLE, [&](fe *st) { schnorr_init(B->sigs[t], st); },
                      [&](size_t step, fe *st) { schnorr_step(step, B->msgs[t], B->sigs[t].s, hbits, st); });
        uint64_t *p = pub + 38 * t;
        for (int j = 0; j < 28; j++) p[j] = f63::from_mont(B->msgs[t][j]);
        for (int j = 0; j < 6; j++) p[28 + j] = f63::from_mont(B->sigs[t].rx[j]);
        for (int j = 0; j < 4; j++) p[34 + j] = B->sigs[t].s.w[j];
    }
    return CSG_OK;
}

}  // extern "C"
