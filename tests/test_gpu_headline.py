"""GPU parity at the HEADLINE sizes (-m gpu): BASELINE.json configs[3] (benches/state_transition.rs shape, 1024 transactions:
trace 2^20 x 94, blowup 8) and its neighbours 512 / 2048 -- the only range (n = 2^19 .. 2^21) that takes `ntt1024_kernel`,
including its inverse / post-scaled / split-extension uses (coset_intt_columns, coset_ntt_entries, the composition
interpolation), which the kernel-level tests do not reach.  Whole proofs are compared byte for byte with the CPU oracle's.
The oracle needs 20-100 s per proof at these sizes (all host cores), so the oracle proofs are computed once per session.

Reference shape: /root/reference/benches/state_transition.rs:13-30 (prove a batch, criterion group over the batch sizes),
options of get_example (/root/reference/src/lib.rs:78-86).
"""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def first_difference(got, want):
    return next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))


@pytest.fixture(scope="module")
def batches(csg):
    """seed -> (trace, pub) of the headline batches, built once"""
    cache = {}

    def get(num_tx, seed=1000):
        key = (num_tx, seed)
        if key not in cache:
            cache[key] = csg.TransactionBatch(seed=seed, num_tx=num_tx).transaction_trace()
        return cache[key]
    return get


@pytest.mark.parametrize("num_tx", [512, 1024, 2048])
def test_headline_proof_identical_to_oracle(ctx, oracle, csg, batches, num_tx):
    trace, pub = batches(num_tx)
    got = ctx.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions())
    want = oracle.prove(oracle.AIR_TRANSACTION, trace, pub, oracle.options())
    assert got == want, f"{num_tx} tx: proof bytes differ (len {len(got)} vs {len(want)}), first difference at byte {first_difference(got, want)}"
    assert csg.verify(csg.AIR_TRANSACTION, pub, got, csg.ProofOptions()) == 0
    if num_tx == 1024:
        # the same trace in the three ways it can cross the boundary: pageable canonical (above), Montgomery words as
        # winterfell's TraceTable stores them, and page-locked memory from csg_host_alloc; then resident re-proving
        mont = oracle.to_mont_fast(trace)
        assert ctx.prove(csg.AIR_TRANSACTION, mont, pub, csg.ProofOptions(), repr=csg.REPR_MONTGOMERY) == want
        hb = csg.HostBuffer(94, trace.shape[1])
        try:
            hb.array[:] = mont
            ctx.set_air(csg.AIR_TRANSACTION, trace.shape[1], pub, csg.ProofOptions())
            assert ctx.prove_trace_ptr(hb.ptr, repr=csg.REPR_MONTGOMERY) == want
            t = ctx.timings()
            assert t["h2d"] > 0, "the overlapped H2D copy must report its own duration"
            ctx.load_trace_ptr(hb.ptr, repr=csg.REPR_MONTGOMERY)
            assert ctx.prove_loaded() == want
            ctx.reload_resident_trace()
            assert ctx.prove_loaded() == want
        finally:
            hb.close()


@pytest.mark.parametrize("ext", [2, 3])
def test_headline_extension_proof_identical_to_oracle(ctx, oracle, csg, batches, ext):
    # the example binary's default is Cubic (/root/reference/examples/state-transition.rs:62-66); Quadratic on half the batch
    num_tx = 1024 if ext == 3 else 512
    trace, pub = batches(num_tx)
    got = ctx.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(field_extension=ext))
    want = oracle.prove_generic(oracle.AIR_TRANSACTION, trace, pub, oracle.options(field_extension=ext))
    assert got == want, f"extension {ext}: first difference at byte {first_difference(got, want)}"
    assert csg.verify(csg.AIR_TRANSACTION, pub, got) == 0


@pytest.mark.parametrize("world", [2, 8])
def test_headline_sharded_proof_is_the_single_gpu_proof(ctx, csg, batches, world):
    # ONE 1024-transaction proof split over `world` contexts by LDE coset (all on this GPU: peer copies stand in for NVLink);
    # every rank must return the bytes of the single-context proof, which the test above pins to the oracle
    trace, pub = batches(1024)
    want = hashlib.sha256(ctx.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions())).hexdigest()
    with csg.LocalGroup(world) as grp:
        proofs = grp.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions())
    assert [hashlib.sha256(p).hexdigest() for p in proofs] == [want] * world


def test_nccl_sharded_check_under_torchrun(csg):
    # the NCCL transport of the sharded proof (one process per GPU): tools/sharded_check.py proves every AIR shape at
    # world = #GPUs and compares each rank's bytes with the single-GPU proof.  Needs at least two GPUs on the box.
    import json
    import subprocess
    import sys
    from pathlib import Path

    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("one GPU on this box: the NCCL transport needs two (the in-process transport is covered above)")
    world = 8 if ngpu >= 8 else 4 if ngpu >= 4 else 2
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(root / "tools" / "sharded_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["world"] == world and res["transport"] == "nccl"
    assert all(c["identical_on_every_rank"] and c["verify"] == 0 for c in res["cases"].values()), res


def test_trace_representations_and_memory_kinds_small(ctx, oracle, csg):
    # the boundary's `repr` argument and the pinned / pageable paths at a size where every column chunk is partial
    batch = csg.TransactionBatch(seed=77, num_tx=2)
    trace, pub = batch.transaction_trace()
    want = oracle.prove(oracle.AIR_TRANSACTION, trace, pub, oracle.options())
    mont = oracle.to_mont_fast(trace)
    assert ctx.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions()) == want
    assert ctx.prove(csg.AIR_TRANSACTION, mont, pub, csg.ProofOptions(), repr=csg.REPR_MONTGOMERY) == want
    # words at or above p are reduced, not trusted: canonical v + p and Montgomery m + p denote the same elements
    hi = trace.copy()
    hi[3, 5] += np.uint64(csg.P)
    assert ctx.prove(csg.AIR_TRANSACTION, hi, pub, csg.ProofOptions()) == want
    mh = mont.copy()
    mh[7, 9] += np.uint64(csg.P)
    assert ctx.prove(csg.AIR_TRANSACTION, mh, pub, csg.ProofOptions(), repr=csg.REPR_MONTGOMERY) == want
    with pytest.raises(csg.CsgError):
        ctx.prove(csg.AIR_TRANSACTION, trace, pub, csg.ProofOptions(), repr=2)
    # one separately allocated array per column, as a winterfell TraceTable holds them (csg_prove_columns)
    columns = [np.array(mont[c]) for c in range(94)]
    assert ctx.prove_columns(csg.AIR_TRANSACTION, columns, pub, csg.ProofOptions(), repr=csg.REPR_MONTGOMERY) == want
    assert ctx.prove_columns(csg.AIR_TRANSACTION, [np.array(trace[c]) for c in range(94)], pub, csg.ProofOptions()) == want
    hb = csg.HostBuffer(94, trace.shape[1])
    try:
        hb.array[:] = trace
        ctx.set_air(csg.AIR_TRANSACTION, trace.shape[1], pub, csg.ProofOptions())
        assert ctx.prove_trace_ptr(hb.ptr) == want
    finally:
        hb.close()
    # csg_host_register: the caller's own allocation, page-locked in place
    import ctypes as C
    own = np.ascontiguousarray(mont)
    assert csg.lib().csg_host_register(C.c_void_p(own.ctypes.data), own.nbytes) == 0
    try:
        assert ctx.prove_trace_ptr(own.ctypes.data, repr=csg.REPR_MONTGOMERY) == want
    finally:
        assert csg.lib().csg_host_unregister(C.c_void_p(own.ctypes.data)) == 0


def test_level2_openings_validate_positions(ctx, csg):
    # caller-supplied query positions were used unchecked (out-of-bounds device reads, u8 truncation)
    import ctypes as C
    trace, pub = csg.build_rescue_trace(np.arange(42, 49, dtype=np.uint64), 32)
    ctx.set_air(csg.AIR_RESCUE, trace.shape[1], pub, csg.ProofOptions(blowup_factor=4))
    ctx.load_trace(trace)
    root = (C.c_uint8 * 32)()
    L = csg.lib()
    assert L.csg_extend_and_commit_trace(ctx._h, root) == 0
    lde_n = trace.shape[1] * 4
    rows, paths, plen = np.zeros(300 * 14, dtype=np.uint64), np.zeros(1 << 20, dtype=np.uint8), C.c_size_t()

    def open_trace(positions):
        pos = np.array(positions, dtype=np.uint64)
        return L.csg_open_trace(ctx._h, pos.ctypes.data_as(C.POINTER(C.c_uint64)), pos.size, rows.ctypes.data_as(C.POINTER(C.c_uint64)),
                                paths.ctypes.data_as(C.POINTER(C.c_uint8)), paths.size, C.byref(plen))
    assert open_trace([0, 5, lde_n - 1]) == 0 and plen.value > 0
    assert open_trace([0, lde_n]) == 1                 # outside the domain
    assert open_trace([3, 9, 3]) == 1                  # duplicate
    assert open_trace(list(range(256))) == 1           # more than a batch opening can describe
    assert open_trace([]) == 1
    assert open_trace([1, 2, 3]) == 0                  # the context stays usable
