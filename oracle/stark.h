/* ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for everything in this header:
 *
 * The STARK engine the reference drives (Prover::prove called at src/lib.rs:140, winterfell::verify at
 * src/lib.rs:149) lives in the un-vendored git dependency
 *     winterfell = { git = "https://github.com/ToposWare/winterfell.git", rev = "8e37310" }   (Cargo.toml:20)
 * which is a fork of facebook/winterfell v0.3.x.  Its source is not available in this environment and there is
 * no Rust toolchain, so this file restates the *published* winterfell v0.3 protocol from memory ([RECALLED] in
 * SURVEY.md Appendix C): domain construction, trace LDE, row hashing + Merkle commitment, constraint merging with
 * degree adjustment, composition polynomial split, OOD frame, DEEP composition, FRI (folding 4) with per-layer
 * commitments, query phase and the StarkProof byte layout.  The reference holds no golden vector for any of it
 * (SURVEY.md section 4), so byte parity with the real fork cannot be claimed; what IS checked is
 *   - the restated verifier (stark_verify) accepts every proof, and rejects tampered proofs / wrong public inputs
 *     exactly like the reference's tests expect (src/tests.rs:12-37);
 *   - the CUDA prover emits byte-identical proofs to this prover.
 */
#ifndef ORACLE_STARK_H
#define ORACLE_STARK_H
#include "air.h"
#include "hashes.h"

typedef struct {
    uint32_t num_queries;       /* 42  (src/lib.rs:79) */
    uint32_t blowup_factor;     /* 8 */
    uint32_t grinding_factor;   /* 0 */
    uint32_t hash_fn;           /* HASH_BLAKE3_256 / HASH_SHA3_256 */
    uint32_t field_extension;   /* 1 = None, 2 = Quadratic, 3 = Cubic (ext.h) */
    uint32_t fri_folding_factor;/* 4 */
    uint32_t fri_max_remainder_size; /* 256 */
} stark_options;

/* intermediate values exposed for stage-level parity tests and debugging */
typedef struct {
    uint8_t trace_root[32], constraint_root[32];
    uint64_t z;                   /* canonical */
    uint32_t num_fri_layers;      /* committed layers incl. the remainder layer */
    uint8_t fri_roots[16][32];
    uint64_t fri_alphas[16];
    uint32_t num_positions;
    uint64_t positions[256];
    uint64_t pow_nonce;
    double t_lde, t_commit_trace, t_constraints, t_composition, t_deep, t_fri, t_queries, t_total; /* seconds */
} stark_debug;

/* trace: column-major, canonical u64, trace[c * n + i].  Returns 0 and a malloc'ed proof. */
int stark_prove(int air_id, const uint64_t *trace, size_t n, const uint64_t *pub, size_t npub, const stark_options *opt,
                uint8_t **proof, size_t *proof_len, stark_debug *dbg);
/* 0 = accepted; negative = malformed proof; positive = the verification step that failed */
int stark_verify(int air_id, const uint64_t *pub, size_t npub, const uint8_t *proof, size_t proof_len);
void stark_free(void *p);
/* the extension-field code path at any degree d = field_extension in {1,2,3}; d = 1 must reproduce stark_prove (self-check) */
int stark_prove_generic(int air_id, const uint64_t *trace, size_t n, const uint64_t *pub, size_t npub, const stark_options *opt, int d,
                        uint8_t **proof, size_t *proof_len);
int stark_verify_generic(int air_id, const uint64_t *pub, size_t npub, const uint8_t *proof, size_t proof_len);
void ext_mul_canonical(int d, const uint64_t *a, const uint64_t *b, uint64_t *out);
void ext_inv_canonical(int d, const uint64_t *a, uint64_t *out);

/* ---- building blocks, exported for kernel-level parity tests ---- */
void ntt_natural(fe *a, size_t n, int inverse);                       /* in place; inverse includes the 1/n scaling */
void lde_column(const fe *evals, size_t n, size_t blowup, fe *out);    /* trace column -> evaluations on offset*<w_lde>, natural order */
void hash_elements(int hash_fn, const fe *elems, size_t n, uint8_t out[32]); /* hash of the canonical LE bytes */
void merkle_build(int hash_fn, const uint8_t *leaves, size_t nleaves, uint8_t *nodes /* 2*nleaves*32 */);
size_t merkle_prove_batch(const uint8_t *nodes, size_t nleaves, const size_t *positions, size_t npos, uint8_t *out /* cap */);
void fri_fold4(const fe *evals, size_t n, fe alpha, fe *out); /* one degree-respecting projection, folding factor 4 */
void fe_array_to_mont(const uint64_t *in, fe *out, size_t n);
void fe_array_from_mont(const fe *in, uint64_t *out, size_t n);
#endif
