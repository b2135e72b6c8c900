import sys; sys.path.insert(0,'.')
import certificate_stark_b200 as csg
for ntx in (512, 1024, 2048):
    b=csg.TransactionBatch(seed=5, num_tx=ntx); pub=b.public_inputs()
    with csg.Context(0) as c:
        c.set_air(csg.AIR_TRANSACTION, 1024*ntx, pub, csg.ProofOptions())
        c.build_transaction_trace(b)
        p=c.prove_loaded()
        print(ntx, len(p), "verify", csg.verify(csg.AIR_TRANSACTION, pub, p), {k:round(v,2) for k,v in c.timings().items() if k in ("lde","constraints","composition","total")}, flush=True)
