"""CPU tests of the product's host side: the C-ABI library loads and exports what include/csg.h declares, the host-compiled
halves (AIR descriptors, fused per-row constraint evaluation, batch-opening shapes, transcript hashes) agree with the oracle,
and nothing pretends to work without the CUDA device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(csg):
    header = (ROOT / "include" / "csg.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(csg_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 35
    L = C.CDLL(str(csg.LIB_PATH))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in csg.h but not exported: {missing}"


def test_no_cpu_fallback(csg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(csg.CsgError, match="no CPU fallback"):
        csg.Context(0)
    with pytest.raises(csg.CsgError):
        csg.get_example(2)
    assert b"no context" in csg.lib().csg_last_error(None)


def test_host_harness_against_oracle(oracle, tmp_path):
    # compiles airs.cuh / air_desc.cpp / transcript.hpp for the host and checks them against the oracle (see host_harness.cpp)
    exe = tmp_path / "host_harness"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", str(ROOT / "tests" / "host_harness.cpp"), "-x", "c++",
                           str(ROOT / "certificate_stark_b200" / "csrc" / "host" / "air_desc.cpp"), f"-L{ROOT / 'oracle'}", "-l:liboracle.so",
                           f"-Wl,-rpath,{ROOT / 'oracle'}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout[-2000:]


def test_witness_builders_reject_bad_shapes(csg):
    with pytest.raises(csg.CsgError):
        csg.build_rescue_trace(np.arange(7, dtype=np.uint64), 3)
    with pytest.raises(csg.CsgError):
        csg.TransactionBatch(seed=1, num_tx=3).transaction_trace()
    with pytest.raises(csg.CsgError):
        csg.TransactionBatch(seed=1, num_tx=0)
    with pytest.raises(csg.CsgError):
        csg.TransactionExample(csg.ProofOptions(), 3)          # src/lib.rs:100-103


def test_witness_is_deterministic_in_the_seed(csg):
    a, pa = csg.TransactionBatch(seed=11, num_tx=2).transaction_trace()
    b, pb = csg.TransactionBatch(seed=11, num_tx=2).transaction_trace()
    c, _ = csg.TransactionBatch(seed=12, num_tx=2).transaction_trace()
    assert np.array_equal(a, b) and np.array_equal(pa, pb) and not np.array_equal(a, c)
    assert a.shape == (94, 2048) and int(a.max()) < csg.P


def test_range_trace_layout(csg):
    # src/range/prover.rs:74-84: column 0 = bits from the most significant one, column 1 = running value
    trace, pub = csg.build_range_trace(0b1011)
    assert int(pub[0]) == 0b1011 and int(trace[1, 63]) == 0b1011 and [int(v) for v in trace[0, 60:64]] == [1, 0, 1, 1]
    assert int(trace[1, 0]) == 0


def test_proof_options_mirror(csg):
    o = csg.ProofOptions()
    assert (o.num_queries, o.blowup_factor, o.grinding_factor, o.hash_fn, o.field_extension, o.fri_folding_factor, o.fri_max_remainder_size) == \
        (42, 8, 0, csg.HASH_BLAKE3_256, 1, 4, 256)      # get_example, src/lib.rs:78-86
