/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * AIR descriptors: a C restatement of what the reference hands winterfell through `impl Air`:
 * context (width, length, transition degrees), evaluate_transition, assertions, periodic columns.
 *   TransactionAir   src/air.rs:64-189, 194-610
 *   MerkleAir        src/merkle/update/air.rs:42-401
 *   PreMerkleAir     src/merkle/init/air.rs:46-211
 *   SchnorrAir       src/schnorr/air.rs:47-585
 *   RangeProofAir    src/range/air.rs:43-105
 *   RescueAir        benches/rescue.rs:163-268
 * Pinned against a second, independent restatement of the same files (oracle/pyair.py): tests/golden/air_vectors.txt holds result[]
 * vectors, degrees, periodic-column fingerprints and assertions for all six AIRs; tests/host_harness.cpp checks this file against them.
 */
#ifndef ORACLE_AIR_H
#define ORACLE_AIR_H
#include "f63.h"

enum { AIR_TRANSACTION = 0, AIR_MERKLE_UPDATE = 1, AIR_MERKLE_INIT = 2, AIR_SCHNORR = 3, AIR_RANGE = 4, AIR_RESCUE = 5 };

typedef struct { uint32_t base, ncycles, cycles[2]; } air_degree;
typedef struct { uint32_t column; size_t first_step, stride; size_t nvalues; fe *values; } air_assertion;

typedef struct air {
    int id;
    uint32_t width;
    size_t trace_len;
    uint32_t num_constraints;
    air_degree *degrees;
    uint32_t num_periodic;
    fe **periodic;        /* num_periodic columns (Montgomery) */
    size_t *periodic_len; /* cycle length of each */
    uint32_t num_assertions;
    air_assertion *assertions; /* in the order get_assertions() returns them */
    uint64_t *pub_inputs;      /* canonical u64 words; their LE bytes are what PublicInputs::write_into emits */
    size_t num_pub_inputs;
    size_t nsig; /* Schnorr only */
    void (*eval)(const struct air *, const fe *cur, const fe *next, const fe *periodic, fe *result);
} air_t;

/* pub layout: TRANSACTION/MERKLE_UPDATE: initial_root[7] final_root[7]; MERKLE_INIT: s[14] r[14] delta;
 * RANGE: number; RESCUE: seed[7] result[7]; SCHNORR: per signature message[28] Rx[6] s[4 LE words]. All canonical. */
air_t *air_new(int air_id, size_t trace_len, const uint64_t *pub, size_t npub);
void air_free(air_t *a);

/* the composite evaluate_constraints of the transaction AIR, exposed for unit tests (src/air.rs:383-610) */
void air_eval_row(const air_t *a, size_t step, const fe *cur, const fe *next, fe *result);
/* winterfell helpers [RECALLED]: evaluation degree of a constraint and ce blowup of the AIR */
size_t air_eval_degree(const air_degree *d, size_t trace_len);
size_t air_ce_blowup(const air_t *a);
#endif
