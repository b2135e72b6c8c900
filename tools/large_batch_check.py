import sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, certificate_stark_b200 as csg
for ntx in (2048, 4096):
    t=time.perf_counter(); b=csg.TransactionBatch(seed=77, num_tx=ntx); pub=b.public_inputs(); print(ntx, "batch %.1fs"%(time.perf_counter()-t), flush=True)
    with csg.Context(0) as c:
        c.set_air(csg.AIR_TRANSACTION, 1024*ntx, pub, csg.ProofOptions())
        c.build_transaction_trace(b)
        p=c.prove_loaded()
        t=time.perf_counter(); c.build_transaction_trace(b); p2=c.prove_loaded(); dt=time.perf_counter()-t
        assert p==p2
        tm=c.timings()
        print(ntx, "witness+prove %.1f ms"%(dt*1e3), {k:round(v,1) for k,v in tm.items() if k in ("h2d","lde","commit_trace","constraints","composition","ood_deep","fri","total")}, "proof", len(p), "verify", csg.verify(csg.AIR_TRANSACTION, pub, p), flush=True)
        # cross-check against the host witness path for the smaller one
        if ntx == 2048:
            tr,_=b.transaction_trace()
            assert c.prove(csg.AIR_TRANSACTION, tr, pub, csg.ProofOptions()) == p
            print("host-witness proof identical")
