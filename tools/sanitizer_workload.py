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
