"""GPU tests (-m gpu) of FieldExtension::Quadratic / Cubic (the reference's own tests prove and verify with
all three settings, src/tests.rs:12-30, src/schnorr/tests.rs, src/merkle/*/tests.rs, src/range/tests.rs): the CUDA prover
must emit the oracle's proof bytes, and both verifiers (the oracle's and the product's host verifier, written independently)
must accept them and reject wrong public inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 0x4180000000000001


def cases(csg):
    z = np.zeros(14, dtype=np.uint64)
    yield "range", csg.AIR_RANGE, *csg.build_range_trace(2**63 - 1), 8
    yield "rescue", csg.AIR_RESCUE, *csg.build_rescue_trace(np.arange(42, 49, dtype=np.uint64), 128), 4
    yield "merkle_init", csg.AIR_MERKLE_INIT, *csg.build_merkle_init_trace(z, z, 1), 4
    batch = csg.TransactionBatch(seed=2, num_tx=2)
    yield "merkle_update", csg.AIR_MERKLE_UPDATE, *batch.merkle_update_trace(), 8
    yield "transaction", csg.AIR_TRANSACTION, *batch.transaction_trace(), 8
    yield "schnorr", csg.AIR_SCHNORR, *csg.SignatureBatch(seed=2, num_sig=2).schnorr_trace(), 8


@pytest.mark.parametrize("ext", [2, 3])
def test_extension_proofs_identical_to_oracle(ctx, oracle, csg, ext):
    for name, air, trace, pub, blowup in cases(csg):
        want = oracle.prove(air, trace, pub, oracle.options(blowup=blowup, field_extension=ext))
        got = ctx.prove(air, trace, pub, csg.ProofOptions(blowup_factor=blowup, field_extension=ext))
        if got != want:
            first = next(i for i, (a, b) in enumerate(zip(got, want)) if a != b)
            raise AssertionError(f"{name}, extension degree {ext}: proofs differ from byte {first} of {len(want)} (GPU proof {len(got)} bytes)")
        assert oracle.verify(air, pub, got) == 0 and csg.verify(air, pub, got) == 0, name
        wrong = pub.copy()
        wrong[-1] = (int(wrong[-1]) + 1) % P
        assert oracle.verify(air, wrong, got) != 0 and csg.verify(air, wrong, got) != 0, name     # src/tests.rs:32-37


@pytest.mark.parametrize("ext,hash_fn,num_tx", [(3, 2, 4), (2, 3, 4), (3, 2, 16)])
def test_example_default_is_cubic(ctx, oracle, csg, ext, hash_fn, num_tx):
    # examples/state-transition.rs:58-71: 4 transactions, cubic extension by default, Blake3 or SHA3
    trace, pub = csg.TransactionBatch(seed=1, num_tx=num_tx).transaction_trace()
    want = oracle.prove(oracle.AIR_TRANSACTION, trace, pub, oracle.options(hash_fn=hash_fn, field_extension=ext))
    got = ctx.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(hash_fn=hash_fn, field_extension=ext))
    assert got == want
    assert csg.verify(csg.AIR_TRANSACTION, pub, got) == 0


@pytest.mark.parametrize("ext", [2, 3])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_extension_proof_sharded(csg, ext, world):
    trace, pub = csg.TransactionBatch(seed=5, num_tx=2).transaction_trace()
    opt = csg.ProofOptions(field_extension=ext)
    with csg.Context(0) as one:
        want = one.prove(csg.AIR_TRANSACTION, trace, pub, opt)
    with csg.LocalGroup(world) as grp:
        assert all(p == want for p in grp.prove(csg.AIR_TRANSACTION, trace, pub, opt))
    trace, pub = csg.TransactionBatch(seed=5, num_tx=2).merkle_update_trace()     # ce blowup 4: idle ranks at world 8
    with csg.Context(0) as one:
        want = one.prove(csg.AIR_MERKLE_UPDATE, trace, pub, opt)
    with csg.LocalGroup(world) as grp:
        assert all(p == want for p in grp.prove(csg.AIR_MERKLE_UPDATE, trace, pub, opt))


def test_large_batch_cubic_verifies(csg):
    # size-independent check at a size the CPU oracle would need minutes for: 256 transactions, cubic extension
    batch = csg.TransactionBatch(seed=3, num_tx=256)
    trace, pub = batch.transaction_trace()
    with csg.Context(0) as c:
        proof = c.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(field_extension=3))
        t = c.timings()
    assert csg.verify(csg.AIR_TRANSACTION, pub, proof) == 0
    assert t["kernel_launches"] > 0
