import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, certificate_stark_b200 as csg
with csg.Context(0) as ctx:
    b = csg.TransactionBatch(seed=1, num_tx=1); tr, pub = b.transaction_trace()
    p1 = ctx.prove(csg.AIR_TRANSACTION, tr, pub, csg.ProofOptions())
    ctx.set_air(csg.AIR_TRANSACTION, 1024, pub, csg.ProofOptions()); ctx.build_transaction_trace(b); p2 = ctx.prove_loaded()
    assert p1 == p2
    s = csg.SignatureBatch(seed=2, num_sig=2); tr, pub = s.schnorr_trace()
    ctx.prove(csg.AIR_SCHNORR, tr, pub, csg.ProofOptions(hash_fn=3))
    tr, pub = csg.build_rescue_trace(np.arange(42, 49, dtype=np.uint64), 8)
    ctx.prove(csg.AIR_RESCUE, tr, pub, csg.ProofOptions(blowup_factor=4))
    tr, pub = csg.build_range_trace(77)
    ctx.prove(csg.AIR_RANGE, tr, pub, csg.ProofOptions())
    tr, pub = csg.TransactionBatch(seed=3, num_tx=2).merkle_update_trace()
    ctx.prove(csg.AIR_MERKLE_UPDATE, tr, pub, csg.ProofOptions())
    print("sanitizer workload done")
    # extension fields, the sharded path (4 ranks on this device) and the warp-per-lane 1024-point NTT passes (n = 2^19: pass A)
    b = csg.TransactionBatch(seed=4, num_tx=2); tr, pub = b.transaction_trace()
    for ext in (2, 3):
        assert csg.verify(csg.AIR_TRANSACTION, pub, ctx.prove(csg.AIR_TRANSACTION, tr, pub, csg.ProofOptions(field_extension=ext))) == 0
    with csg.LocalGroup(4) as grp:
        ps = grp.prove(csg.AIR_TRANSACTION, tr, pub, csg.ProofOptions(field_extension=3))
        assert all(p == ps[0] for p in ps)
    rng = np.random.default_rng(1)
    cols = (rng.integers(0, 2**63, size=(2, 1 << 19), dtype=np.uint64) % np.uint64(csg.P)).astype(np.uint64)
    ctx.lde(cols, 2)
    print("sanitizer workload (extension, sharded, 1024-point passes) done")
