"""GPU tests (-m gpu) of the device-side batch builder (TransactionMetadata::build_random,
/root/reference/src/lib.rs:235-464): the account tree's whole update history, the authentication paths, the signatures and the
packed witness records are computed by kernels from a host plan that hashes nothing.  The result must be bit-identical to the
host builder (csg_tx_batch_new + csg_tx_batch_pack, witness.cpp), whose witnesses the parity tests pin to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,num_tx,depth", [(1, 1, 15), (7, 2, 15), (3, 16, 15), (11, 64, 15), (5, 4, 3), (9, 32, 7)])
def test_device_batch_records_are_the_host_builders(ctx, csg, seed, num_tx, depth):
    # depth 3 is the reference's cfg(test) tree (src/merkle/constants.rs:21-25): few slots, so accounts repeat and senders collide
    want = csg.TransactionBatch(seed=seed, num_tx=num_tx, tree_depth=depth)
    pub = ctx.build_batch(seed, num_tx, depth)
    assert np.array_equal(pub, want.public_inputs())
    got, ref = ctx.download_batch_records(num_tx), want.packed_records()
    bad = np.argwhere(got != ref)
    assert bad.size == 0, f"first differing (transfer, word): {bad[0]} (layout: certificate_stark_b200/csrc/witness.cuh)"
    assert ctx.timings()["batch_build"] > 0


def test_metadata_witness_and_proof_all_on_the_device(ctx, oracle, csg):
    # TransactionExample::new + prove (src/lib.rs:92-141) with nothing but seeds and the plan crossing PCIe
    num_tx, seed = 8, 21
    host = csg.TransactionBatch(seed=seed, num_tx=num_tx)
    trace, pub = host.transaction_trace()
    got_pub = ctx.build_batch(seed, num_tx)
    assert np.array_equal(got_pub, pub)
    ctx.set_air(csg.AIR_TRANSACTION, 1024 * num_tx, got_pub, csg.ProofOptions())
    ctx.build_transaction_trace_resident()
    assert np.array_equal(ctx.download_trace(94, 1024 * num_tx), trace)
    proof = ctx.prove_loaded()
    assert proof == oracle.prove(oracle.AIR_TRANSACTION, trace, pub, oracle.options())
    ex = csg.TransactionExample(csg.ProofOptions(), num_tx, seed=seed, batch_on_device=True)
    assert ex.prove() == proof and ex.verify(proof) and not ex.verify_with_wrong_inputs(proof)


def test_device_batch_errors(ctx, csg):
    with pytest.raises(csg.CsgError):
        ctx.build_batch(1, 0)
    with pytest.raises(csg.CsgError):
        ctx.build_batch(1, 4, 5)            # depth + 1 must be a power of two (src/lib.rs:106-109)
    with csg.Context(0) as fresh:
        with pytest.raises(csg.CsgError):
            fresh.download_batch_records(1)
        pub = fresh.build_batch(2, 2)
        fresh.set_air(csg.AIR_TRANSACTION, 1024, pub, csg.ProofOptions())
        with pytest.raises(csg.CsgError):   # the resident batch has 2 transfers, the AIR was set for 1
            fresh.build_transaction_trace_resident()


@pytest.mark.parametrize("num_tx", [1, 4, 32])
def test_device_merkle_update_witness_matches_host_builder(csg, oracle, num_tx):
    # SURVEY.md 8(f).1 for MerkleProver::build_trace (src/merkle/update/prover.rs:37-80), bit tweak at step 1 included
    batch = csg.TransactionBatch(seed=31 + num_tx, num_tx=num_tx)
    want, pub = batch.merkle_update_trace()
    with csg.Context(0) as c:
        c.set_air(csg.AIR_MERKLE_UPDATE, 512 * num_tx, pub, csg.ProofOptions())
        c.build_merkle_update_trace(batch)
        got = c.download_trace(65, 512 * num_tx)
        bad = np.argwhere(got != want)
        assert bad.size == 0, f"first differing (column, row): {bad[0]}"
        proof = c.prove_loaded()
    assert proof == oracle.prove(oracle.AIR_MERKLE_UPDATE, want, pub, oracle.options())


@pytest.mark.parametrize("num_sig", [1, 2, 16])
def test_device_schnorr_witness_matches_host_builder(csg, oracle, num_sig):
    # SURVEY.md 8(f).1 for SchnorrProver::build_trace (src/schnorr/prover.rs:52-80): messages with 16 random elements
    batch = csg.SignatureBatch(seed=41 + num_sig, num_sig=num_sig)
    want, pub = batch.schnorr_trace()
    with csg.Context(0) as c:
        c.set_air(csg.AIR_SCHNORR, 512 * num_sig, pub, csg.ProofOptions())
        c.build_schnorr_trace(batch)
        got = c.download_trace(56, 512 * num_sig)
        bad = np.argwhere(got != want)
        assert bad.size == 0, f"first differing (column, row): {bad[0]}"
        proof = c.prove_loaded()
        with pytest.raises(csg.CsgError):       # a batch of another size
            c.build_schnorr_trace(csg.SignatureBatch(seed=1, num_sig=num_sig * 2))
    assert proof == oracle.prove(oracle.AIR_SCHNORR, want, pub, oracle.options())
