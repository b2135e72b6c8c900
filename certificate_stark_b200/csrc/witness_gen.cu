// Witness generation on the GPU: TransactionProver::build_trace of the reference
// (/root/reference/src/prover.rs:37-98 with src/trace.rs:28-142, src/merkle/update/trace.rs:19-136,
// src/schnorr/trace.rs:18-122, src/range/prover.rs:65-84) -- the 1024-row fragment of every transaction, written as the
// canonical column-major table the proving stages consume, without the trace ever existing in host memory.
//
// A fragment is a handful of independent sequential chains: four Rescue hash states climbing the two Merkle paths
// (127 rounds each), the Rescue state hashing the signed message (40 steps), two double-and-add scalar multiplications
// (510 steps each) and the final point addition, plus columns that are constant or simple bit accumulators.  Each chain
// kind has its own kernel with ONE THREAD PER (transaction, chain): the 32 lanes of a warp run the same code on 32
// different transactions.  The chains are latency-bound (a few ms) and leave the GPU mostly idle, which is fine: the host
// builder this replaces needs seconds for the same batch.
#include "ecc.cuh"
#include "rescue.cuh"
#include "witness.cuh"

namespace csg {
using namespace f63;

namespace {
enum : int {
    HSW = 14, HRW = 7, APW = 12, PPW = 18, TX_ROWS = 1024, MERKLE_ROWS = 512, SCALAR_MUL_LENGTH = 510, NUM_HASH_ITER = 5, RANGE_LOG = 64,
    SENDER_KEY = 65, DELTA_COPY = 89, PREV_ROOT = 58, SIG_HASH = 42, LIMBS = 37, TX_WIDTH = 94
};

struct Out {   // canonical column-major trace; item t (a transfer or a signature) owns rows t*item_rows .. (t+1)*item_rows - 1
    uint64_t *trace;
    unsigned long long n;
    unsigned item_rows;      // 1024: a transfer of the transaction AIR; 512: a fragment of a standalone sub-AIR
    unsigned schnorr_row0;   // first row of the signature phase inside an item (512 in a transfer, 0 standalone)
    __device__ __forceinline__ void put(unsigned col, unsigned long long row, fe v) const { trace[col * n + row] = from_mont(v); }
};
// word m of the signed message of a record (witness.cuh)
__device__ __forceinline__ fe message_word(const uint64_t *T, unsigned m) {
    return m < APW ? T[WIT_S_OLD + m] : m < 2 * APW ? T[WIT_R_OLD + m - APW] : m == 2 * APW ? T[WIT_DELTA] : m == 2 * APW + 1 ? T[WIT_S_OLD + APW + 1] : T[WIT_M26 + m - 26];
}
__device__ __forceinline__ bool bit256(const uint64_t *w, unsigned i) { return (w[i >> 6] >> (i & 63)) & 1; }

// ---- the four Merkle-path hash states of a transaction: thread = (transaction, state), rows 0..511 of its 14 columns
__global__ void __launch_bounds__(64) wit_merkle_kernel(const uint64_t *__restrict__ in, unsigned ntx, unsigned depth, Out out) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ntx * 4) return;
    const unsigned tx = g >> 2, s = g & 3;
    const uint64_t *T = in + (size_t)tx * WIT_WORDS;
    const bool receiver = s >= 2, updated = s & 1;
    const unsigned col0 = 15 * s - (s >> 1);
    const unsigned long long row0 = (unsigned long long)tx * out.item_rows;
    const uint64_t *acct = T + (receiver ? WIT_R_OLD : WIT_S_OLD), *path = T + (receiver ? WIT_R_PATH : WIT_S_PATH);
    const uint64_t index = T[receiver ? WIT_R_IDX : WIT_S_IDX];
    const fe delta = T[WIT_DELTA];
    fe st[14];
    for (int i = 0; i < 14; i++) st[i] = acct[i];
    if (updated) {   // src/merkle/update/trace.rs:19-48
        if (!receiver) { st[APW] = sub(st[APW], delta); st[APW + 1] = add(st[APW + 1], ONE); }
        else st[APW] = add(st[APW], delta);
    }
    const unsigned hash_len = 8 * depth + 7;
    fe bit = 0;
    const bool owns_bit = !updated;            // columns 14 and 43 travel with the initial states
    const bool owns_root = s == 3;             // columns 58..64: previous root, replaced by the new one after the last round
    for (int i = 0; i < 14; i++) out.put(col0 + i, row0, st[i]);
    if (owns_bit) out.put(col0 + HSW, row0, 0);
    if (owns_root) for (int i = 0; i < HRW; i++) out.put(PREV_ROOT + i, row0, T[WIT_ROOT + i]);
    fe root[7];
    for (int i = 0; i < HRW; i++) root[i] = T[WIT_ROOT + i];
    for (unsigned step = 0; step + 1 < MERKLE_ROWS; step++) {
        if (step < hash_len) {   // src/merkle/update/trace.rs:97-136
            if (step % 8 < 7) rescue::apply_round(st, step);
            else {
                const uint64_t *node = path + (step / 8 + 1) * HRW;
                const bool b = (index >> (step / 8)) & 1;
                for (int i = 0; i < HRW; i++) {
                    if (!b) st[HRW + i] = node[i];
                    else { st[HRW + i] = st[i]; st[i] = node[i]; }
                }
                bit = b ? ONE : 0;
            }
            if (owns_root && step == hash_len - 1) for (int i = 0; i < HRW; i++) root[i] = st[i];
        }
        const unsigned long long row = row0 + step + 1;
        for (int i = 0; i < 14; i++) out.put(col0 + i, row, st[i]);
        if (owns_bit) out.put(col0 + HSW, row, bit);
        if (owns_root) for (int i = 0; i < HRW; i++) out.put(PREV_ROOT + i, row, root[i]);
    }
    if (owns_root)   // the root register is not touched by the Schnorr half of the fragment (src/trace.rs:89-100)
        for (unsigned r = MERKLE_ROWS; r < out.item_rows; r++) for (int i = 0; i < HRW; i++) out.put(PREV_ROOT + i, row0 + r, root[i]);
}

// ---- Rescue state hashing R.x and the message (columns 42..55, rows 512..1023): thread = transaction
__global__ void __launch_bounds__(64) wit_sig_hash_kernel(const uint64_t *__restrict__ in, unsigned ntx, Out out) {
    const unsigned tx = blockIdx.x * blockDim.x + threadIdx.x;
    if (tx >= ntx) return;
    const uint64_t *T = in + (size_t)tx * WIT_WORDS;
    const unsigned long long row0 = (unsigned long long)tx * out.item_rows + out.schnorr_row0;
    fe st[14];
    for (int i = 0; i < 14; i++) st[i] = i < 6 ? T[WIT_RX + i] : 0;      // src/schnorr/trace.rs:18-30
    for (int i = 0; i < 14; i++) out.put(SIG_HASH + i, row0, st[i]);
    for (unsigned ss = 0; ss + 1 < MERKLE_ROWS; ss++) {                    // src/schnorr/trace.rs:43-67
        if (ss < 8 * NUM_HASH_ITER) {
            if (ss % 8 < 7) rescue::apply_round(st, ss);
            else if (ss < 8 * (NUM_HASH_ITER - 1)) {
                for (int i = 0; i < HRW; i++) st[HRW + i] = message_word(T, HRW * (ss / 8) + i);   // sender key | receiver key | delta | nonce | 0 | 0 for a transfer (src/lib.rs:467-481)
            } else for (int i = 0; i < HRW; i++) st[HRW + i] = 0;
        }
        for (int i = 0; i < 14; i++) out.put(SIG_HASH + i, row0 + ss + 1, st[i]);
    }
}

__device__ __forceinline__ void put_point(const Out &out, unsigned col0, unsigned long long row, const ecc::point &p) {
    for (int i = 0; i < 6; i++) { out.put(col0 + i, row, p.x.c[i]); out.put(col0 + 6 + i, row, p.y.c[i]); out.put(col0 + 12 + i, row, p.z.c[i]); }
}
// ---- one scalar multiplication (bank 0: s.G in columns 0..18; bank 1: h.P in 19..37 with the limbs of h in 38..41),
// rows 512..1022; the final states are kept for wit_final_kernel.  thread = (transaction, bank)
__global__ void __launch_bounds__(64) wit_scalar_mult_kernel(const uint64_t *__restrict__ in, unsigned ntx, Out out, fe *__restrict__ finals) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ntx * 2) return;
    const unsigned tx = g >> 1, bank = g & 1, col0 = bank * (PPW + 1);
    const uint64_t *T = in + (size_t)tx * WIT_WORDS;
    const uint64_t *bits = T + (bank ? WIT_H : WIT_S);
    const unsigned long long row0 = (unsigned long long)tx * out.item_rows + out.schnorr_row0;
    ecc::fp6 qx, qy;
    {
        const uint64_t *q = bank ? T + WIT_S_OLD : CSG_TABLE(CSG_GENERATOR);   // h multiplies the sender's public key
        for (int i = 0; i < 6; i++) { qx.c[i] = q[i]; qy.c[i] = q[6 + i]; }
    }
    ecc::point p{};
    p.y.c[0] = ONE;   // the identity (0 : 1 : 0)
    fe limb[4] = {0, 0, 0, 0}, bit = 0;
    put_point(out, col0, row0, p);
    out.put(col0 + PPW, row0, 0);
    if (bank) for (int i = 0; i < 4; i++) out.put(LIMBS + 1 + i, row0, 0);
    for (unsigned ss = 0; ss < SCALAR_MUL_LENGTH; ss++) {   // src/schnorr/trace.rs:70-104
        const unsigned real = ss / 2, chunk = real < 63 ? 0 : (real - 63) / 64 + 1;
        const bool b = bit256(bits, 254 - real);
        bit = b ? ONE : 0;
        if (ss % 2 == 0) {
            p = ecc::double_point(p);
            if (bank) limb[3 - chunk] = add(dbl(limb[3 - chunk]), bit);
        } else if (b) p = ecc::add_mixed(p, qx, qy);
        const unsigned long long row = row0 + ss + 1;
        put_point(out, col0, row, p);
        out.put(col0 + PPW, row, bit);
        if (bank) for (int i = 0; i < 4; i++) out.put(LIMBS + 1 + i, row, limb[i]);
    }
    fe *F = finals + ((size_t)tx * 2 + bank) * 24;
    for (int i = 0; i < 6; i++) { F[i] = p.x.c[i]; F[6 + i] = p.y.c[i]; F[12 + i] = p.z.c[i]; }
    F[18] = bit;
    for (int i = 0; i < 4; i++) F[19 + i] = limb[i];
}

// ---- last row of the signature (S + h.P with x reduced to affine) and every column that is a copy or a bit accumulator:
// 56/57 and 92/93 (range proofs of delta and sigma), 65..91 (keys, delta, sigma, nonce).  thread = transaction
__global__ void __launch_bounds__(64) wit_final_kernel(const uint64_t *__restrict__ in, unsigned ntx, Out out, const fe *__restrict__ finals) {
    const unsigned tx = blockIdx.x * blockDim.x + threadIdx.x;
    if (tx >= ntx) return;
    const uint64_t *T = in + (size_t)tx * WIT_WORDS;
    const unsigned long long row0 = (unsigned long long)tx * out.item_rows, last = row0 + out.item_rows - 1;
    {   // src/schnorr/trace.rs:105-119
        const fe *A = finals + (size_t)tx * 48, *B = A + 24;
        ecc::point s, hp;
        for (int i = 0; i < 6; i++) { s.x.c[i] = A[i]; s.y.c[i] = A[6 + i]; s.z.c[i] = A[12 + i]; hp.x.c[i] = B[i]; hp.y.c[i] = B[6 + i]; hp.z.c[i] = B[12 + i]; }
        ecc::point r = ecc::add_full(s, hp);
        ecc::fp6 x = ecc::mul(r.x, ecc::inv(r.z));
        for (int i = 0; i < 6; i++) { out.put(i, last, x.c[i]); out.put(6 + i, last, r.y.c[i]); out.put(12 + i, last, r.z.c[i]); }
        out.put(PPW, last, ONE);
        for (int i = 0; i < 18; i++) out.put(PPW + 1 + i, last, B[i]);
        out.put(LIMBS, last, B[18]);
        for (int i = 0; i < 4; i++) out.put(LIMBS + 1 + i, last, B[19 + i]);
    }
    if (out.item_rows != TX_ROWS) return;   // a standalone signature has no copy or range-proof columns
    const fe delta = T[WIT_DELTA], sigma = sub(T[WIT_S_OLD + APW], delta), nonce = T[WIT_S_OLD + APW + 1];
    const uint64_t dbits = from_mont(delta), sbits = from_mont(sigma);
    fe dacc = 0, sacc = 0, dbit = 0, sbit = 0;
    for (unsigned r = 0; r < TX_ROWS; r++) {
        const unsigned long long row = row0 + r;
        for (int i = 0; i < APW; i++) { out.put(SENDER_KEY + i, row, T[WIT_S_OLD + i]); out.put(SENDER_KEY + APW + i, row, T[WIT_R_OLD + i]); }   // src/trace.rs:28-53
        out.put(DELTA_COPY, row, delta); out.put(DELTA_COPY + 1, row, sigma); out.put(DELTA_COPY + 2, row, nonce);
        if (r > MERKLE_ROWS && r <= MERKLE_ROWS + RANGE_LOG) {   // src/trace.rs:113-128, src/range/prover.rs:74-84
            const unsigned ss = r - MERKLE_ROWS - 1;
            dbit = ((dbits >> (RANGE_LOG - 1 - ss)) & 1) ? ONE : 0; dacc = add(dbl(dacc), dbit);
            sbit = ((sbits >> (RANGE_LOG - 1 - ss)) & 1) ? ONE : 0; sacc = add(dbl(sacc), sbit);
        }
        out.put(92, row, sbit); out.put(93, row, sacc);
        if (r >= MERKLE_ROWS) { out.put(56, row, dbit); out.put(57, row, dacc); }   // rows below 512 of 56/57 belong to a hash state
    }
}

}  // namespace

__global__ void wit_set_one_kernel(uint64_t *a, uint64_t *b) { *a = 1; *b = 1; }

void build_merkle_update_trace(const uint64_t *inputs_dev, size_t ntx, unsigned tree_depth, uint64_t *trace_dev, Stream &st) {
    const unsigned long long n = (unsigned long long)ntx * MERKLE_ROWS;
    Out out{trace_dev, n, MERKLE_ROWS, 0};
    const unsigned T = 64;
    CSG_LAUNCH(st, wit_merkle_kernel, (unsigned)((ntx * 4 + T - 1) / T), T, 0, inputs_dev, (unsigned)ntx, tree_depth, out);
    // trace.set(SENDER_BIT_POS, 1, ONE); trace.set(RECEIVER_BIT_POS, 1, ONE)   (src/merkle/update/prover.rs:72-77; canonical 1)
    CSG_LAUNCH(st, wit_set_one_kernel, 1, 1, 0, trace_dev + 14 * n + 1, trace_dev + 43 * n + 1);
}
void build_schnorr_trace(const uint64_t *inputs_dev, size_t nsig, uint64_t *trace_dev, fe *finals_dev, Stream &st) {
    Out out{trace_dev, (unsigned long long)nsig * MERKLE_ROWS, MERKLE_ROWS, 0};
    const unsigned T = 64;
    CSG_LAUNCH(st, wit_sig_hash_kernel, (unsigned)((nsig + T - 1) / T), T, 0, inputs_dev, (unsigned)nsig, out);
    CSG_LAUNCH(st, wit_scalar_mult_kernel, (unsigned)((nsig * 2 + T - 1) / T), T, 0, inputs_dev, (unsigned)nsig, out, finals_dev);
    CSG_LAUNCH(st, wit_final_kernel, (unsigned)((nsig + T - 1) / T), T, 0, inputs_dev, (unsigned)nsig, out, (const fe *)finals_dev);
}

void build_transaction_trace(const uint64_t *inputs_dev, size_t ntx, unsigned tree_depth, uint64_t *trace_dev, fe *finals_dev, Stream &st) {
    Out out{trace_dev, (unsigned long long)ntx * TX_ROWS, TX_ROWS, MERKLE_ROWS};
    const unsigned T = 64;
    CSG_LAUNCH(st, wit_merkle_kernel, (unsigned)((ntx * 4 + T - 1) / T), T, 0, inputs_dev, (unsigned)ntx, tree_depth, out);
    CSG_LAUNCH(st, wit_sig_hash_kernel, (unsigned)((ntx + T - 1) / T), T, 0, inputs_dev, (unsigned)ntx, out);
    CSG_LAUNCH(st, wit_scalar_mult_kernel, (unsigned)((ntx * 2 + T - 1) / T), T, 0, inputs_dev, (unsigned)ntx, out, finals_dev);
    CSG_LAUNCH(st, wit_final_kernel, (unsigned)((ntx + T - 1) / T), T, 0, inputs_dev, (unsigned)ntx, out, (const fe *)finals_dev);
}

}  // namespace csg
