"""GPU tests (-m gpu) of the coset-sharded proof (csg_dist_* of include/csg.h): one proof split over G
contexts by LDE coset must be byte-identical to the single-context proof -- which the parity tests pin to the CPU oracle --
on every rank.  The ranks run as threads of this process over the in-process transport, all on cuda:0, so a one-GPU box
covers the whole sharded code path (ownership, exchanges, interleaving, row sums); the NCCL transport is covered by
tools/sharded_check.py under torchrun on a multi-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sharded_equals_single(csg, oracle, air, trace, pub, options, world):
    with csg.Context(0) as one:
        want = one.prove(air, trace, pub, options)
    with csg.LocalGroup(world) as grp:
        assert [c.dist_info() for c in grp.ctxs] == [(r, world) for r in range(world)]
        proofs = grp.prove(air, trace, pub, options)
        launches = [c.timings()["kernel_launches"] for c in grp.ctxs]
    assert all(p == want for p in proofs), f"sharded proof differs from the single-GPU proof at world={world}"
    assert all(l > 0 for l in launches)
    assert csg.verify(air, pub, want) == 0
    return want


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("num_tx", [1, 4])
def test_transaction_proof_sharded(csg, oracle, world, num_tx):
    trace, pub = csg.TransactionBatch(seed=7, num_tx=num_tx).transaction_trace()
    proof = sharded_equals_single(csg, oracle, csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(), world)
    if num_tx == 1:
        assert proof == oracle.prove(oracle.AIR_TRANSACTION, trace, pub, oracle.options())


@pytest.mark.parametrize("world", [2, 4, 8])
def test_merkle_update_sharded_with_fewer_ce_cosets_than_ranks(csg, oracle, world):
    # ce blowup 4 < blowup 8: at world 8 half of the ranks own no constraint-evaluation coset
    trace, pub = csg.TransactionBatch(seed=3, num_tx=2).merkle_update_trace()
    sharded_equals_single(csg, oracle, csg.AIR_MERKLE_UPDATE, trace, pub, csg.ProofOptions(), world)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("hash_fn", [2, 3])
def test_schnorr_sharded(csg, oracle, world, hash_fn):
    trace, pub = csg.SignatureBatch(seed=5, num_sig=2).schnorr_trace()
    sharded_equals_single(csg, oracle, csg.AIR_SCHNORR, trace, pub, csg.ProofOptions(hash_fn=hash_fn), world)


@pytest.mark.parametrize("world", [2, 4])
def test_rescue_blowup4_sharded(csg, oracle, world):
    trace, pub = csg.build_rescue_trace(np.arange(42, 49, dtype=np.uint64), 128)
    sharded_equals_single(csg, oracle, csg.AIR_RESCUE, trace, pub, csg.ProofOptions(blowup_factor=4), world)


def test_range_two_columns_over_eight_ranks(csg, oracle):
    # 2 columns over 8 ranks: most ranks interpolate nothing; 64-row trace
    trace, pub = csg.build_range_trace(123456789)
    sharded_equals_single(csg, oracle, csg.AIR_RANGE, trace, pub, csg.ProofOptions(), 8)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("ext", [1, 3])
def test_split_exchange_between_ranks_on_a_small_trace(csg, oracle, monkeypatch, world, ext):
    # ranks that own whole even/odd coset pairs keep the low-degree split and exchange the interpolants; by default only traces
    # of 2^14 rows or more take that path (the headline test), CSG_SPLIT_MIN_ROWS=0 forces it here
    trace, pub = csg.TransactionBatch(seed=13, num_tx=2).transaction_trace()
    with csg.Context(0) as one:
        want = one.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(field_extension=ext))
    monkeypatch.setenv("CSG_SPLIT_MIN_ROWS", "0")
    assert sharded_equals_single(csg, oracle, csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(field_extension=ext), world) == want


def test_world_must_divide_blowup(csg):
    trace, pub = csg.build_rescue_trace(np.arange(42, 49, dtype=np.uint64), 128)
    with csg.LocalGroup(8) as grp:
        with pytest.raises(csg.CsgError):
            grp.ctxs[0].set_air(csg.AIR_RESCUE, trace.shape[1], pub, csg.ProofOptions(blowup_factor=4))


def test_resident_trace_sharded_reproves(csg):
    # the benchmark path: trace loaded once on every rank, proved twice from HBM
    trace, pub = csg.TransactionBatch(seed=11, num_tx=2).transaction_trace()
    with csg.Context(0) as one:
        want = one.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions())

    def work(ctx, r):
        ctx.set_air(csg.AIR_TRANSACTION, trace.shape[1], pub, csg.ProofOptions())
        ctx.load_trace(trace)
        a = ctx.prove_loaded()
        ctx.reload_resident_trace()
        return a, ctx.prove_loaded(), ctx.timings()
    with csg.LocalGroup(4) as grp:
        out = grp.run(work)
    for a, b2, t in out:
        assert a == want and b2 == want
        assert t["comm"] > 0
