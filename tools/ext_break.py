import sys, pathlib
sys.path.insert(0, '/root/repo')
import certificate_stark_b200 as csg
ntx=1024
b = csg.TransactionBatch(seed=5, num_tx=ntx); pub = b.public_inputs()
with csg.Context(0) as c:
    for ext in (1, 3):
        c.set_air(csg.AIR_TRANSACTION, 1024 * ntx, pub, csg.ProofOptions(field_extension=ext))
        for _ in range(2):
            c.build_transaction_trace(b); p = c.prove_loaded()
        t = c.timings()
        print(ext, {k: round(v, 2) for k, v in t.items() if isinstance(v, float)}, t["stage_launches"], flush=True)
