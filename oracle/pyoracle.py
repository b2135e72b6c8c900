"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
P = 0x4180000000000001

AIR_TRANSACTION, AIR_MERKLE_UPDATE, AIR_MERKLE_INIT, AIR_SCHNORR, AIR_RANGE, AIR_RESCUE = range(6)
HASH_BLAKE3_256, HASH_SHA3_256 = 2, 3


class StarkOptions(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_queries", "blowup_factor", "grinding_factor", "hash_fn", "field_extension",
                                          "fri_folding_factor", "fri_max_remainder_size")]


class StarkDebug(C.Structure):
    _fields_ = [("trace_root", C.c_uint8 * 32), ("constraint_root", C.c_uint8 * 32), ("z", C.c_uint64),
                ("num_fri_layers", C.c_uint32), ("fri_roots", (C.c_uint8 * 32) * 16), ("fri_alphas", C.c_uint64 * 16),
                ("num_positions", C.c_uint32), ("positions", C.c_uint64 * 256), ("pow_nonce", C.c_uint64),
                ("t_lde", C.c_double), ("t_commit_trace", C.c_double), ("t_constraints", C.c_double),
                ("t_composition", C.c_double), ("t_deep", C.c_double), ("t_fri", C.c_double), ("t_queries", C.c_double),
                ("t_total", C.c_double)]


def build(force=False):
    so = HERE / "liboracle.so"
    srcs = list(HERE.glob("*.c")) + list(HERE.glob("*.h"))
    if force or not so.exists() or (all(s.exists() for s in srcs) and so.stat().st_mtime < max(s.stat().st_mtime for s in srcs)):
        subprocess.check_call(["make", "-C", str(HERE), "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        u64p, u8p, szp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
        L = _lib
        L.stark_prove.argtypes = [C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.POINTER(StarkOptions), C.POINTER(u8p), szp,
                                  C.POINTER(StarkDebug)]
        L.stark_prove.restype = C.c_int
        L.stark_verify.argtypes = [C.c_int, u64p, C.c_size_t, u8p, C.c_size_t]
        L.stark_verify.restype = C.c_int
        L.stark_free.argtypes = [C.c_void_p]
        L.stark_prove_generic.argtypes = [C.c_int, u64p, C.c_size_t, u64p, C.c_size_t, C.POINTER(StarkOptions), C.c_int, C.POINTER(u8p), szp]
        L.stark_prove_generic.restype = C.c_int
        L.stark_verify_generic.argtypes = [C.c_int, u64p, C.c_size_t, u8p, C.c_size_t]
        L.stark_verify_generic.restype = C.c_int
        L.ext_mul_canonical.argtypes = [C.c_int, u64p, u64p, u64p]
        L.ext_inv_canonical.argtypes = [C.c_int, u64p, u64p]
        L.blake3_256.argtypes = [C.c_char_p, C.c_size_t, u8p]
        L.sha3_256.argtypes = [C.c_char_p, C.c_size_t, u8p]
        L.ntt_natural.argtypes = [u64p, C.c_size_t, C.c_int]
        L.lde_column.argtypes = [u64p, C.c_size_t, C.c_size_t, u64p]
        L.hash_elements.argtypes = [C.c_int, u64p, C.c_size_t, u8p]
        L.merkle_build.argtypes = [C.c_int, u8p, C.c_size_t, u8p]
        L.fri_fold4.argtypes = [u64p, C.c_size_t, C.c_uint64, u64p]
        L.rescue_apply_round.argtypes = [u64p, C.c_size_t]
        L.rescue_apply_permutation.argtypes = [u64p]
        L.rescue_digest.argtypes = [u64p, C.c_size_t, u64p]
        L.rescue_merge.argtypes = [u64p, u64p, u64p]
        L.air_new.argtypes = [C.c_int, C.c_size_t, u64p, C.c_size_t]
        L.air_new.restype = C.c_void_p
        L.air_free.argtypes = [C.c_void_p]
        L.air_eval_row.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
        L.air_ce_blowup.argtypes = [C.c_void_p]
        L.air_ce_blowup.restype = C.c_size_t
        L.fe_array_to_mont.argtypes = [u64p, u64p, C.c_size_t]
        L.fe_array_from_mont.argtypes = [u64p, u64p, C.c_size_t]
        L.rescue_init_tables()
        L.ecc_init_tables()
    return _lib


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


R = (1 << 64) % P
RINV = pow(R, -1, P)


def to_mont(a):
    """canonical u64 numpy array -> Montgomery form (python ints, fine for test sizes)"""
    return np.array([(int(v) * R) % P for v in np.asarray(a, dtype=np.uint64).ravel()], dtype=np.uint64).reshape(np.shape(a))


def from_mont(a):
    return np.array([(int(v) * RINV) % P for v in np.asarray(a, dtype=np.uint64).ravel()], dtype=np.uint64).reshape(np.shape(a))


def to_mont_fast(a):
    """canonical -> Montgomery through the C oracle (any shape)"""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().fe_array_to_mont(_p64(a), _p64(out), a.size)
    return out


def from_mont_fast(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().fe_array_from_mont(_p64(a), _p64(out), a.size)
    return out


def options(num_queries=42, blowup=8, grinding=0, hash_fn=HASH_BLAKE3_256, field_extension=1, folding=4, max_remainder=256):
    return StarkOptions(num_queries, blowup, grinding, hash_fn, field_extension, folding, max_remainder)


def prove(air_id, trace, pub, opt, want_debug=False):
    """trace: (width, n) canonical uint64, C-contiguous.  Returns proof bytes (and the debug struct)."""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    pub = np.ascontiguousarray(pub, dtype=np.uint64)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    dbg = StarkDebug()
    rc = lib().stark_prove(air_id, _p64(trace), trace.shape[1], _p64(pub), pub.size, C.byref(opt), C.byref(out), C.byref(n), C.byref(dbg))
    if rc != 0:
        raise RuntimeError(f"oracle stark_prove failed: {rc}")
    proof = bytes(C.cast(out, C.POINTER(C.c_uint8 * n.value)).contents)
    lib().stark_free(out)
    return (proof, dbg) if want_debug else proof


def prove_generic(air_id, trace, pub, opt):
    """the extension-field code path of the oracle at degree d = opt.field_extension (1 included: must equal prove())"""
    trace = np.ascontiguousarray(trace, dtype=np.uint64)
    pub = np.ascontiguousarray(pub, dtype=np.uint64)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = lib().stark_prove_generic(air_id, _p64(trace), trace.shape[1], _p64(pub), pub.size, C.byref(opt), int(opt.field_extension), C.byref(out), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle stark_prove_generic failed: {rc}")
    proof = bytes(C.cast(out, C.POINTER(C.c_uint8 * n.value)).contents)
    lib().stark_free(out)
    return proof


def verify_generic(air_id, pub, proof):
    pub = np.ascontiguousarray(pub, dtype=np.uint64)
    buf = np.frombuffer(proof, dtype=np.uint8)
    return lib().stark_verify_generic(air_id, _p64(pub), pub.size, _p8(buf), buf.size)


def ext_mul(d, a, b):
    a, b, out = np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64), np.zeros(d, dtype=np.uint64)
    lib().ext_mul_canonical(d, _p64(a), _p64(b), _p64(out))
    return [int(v) for v in out]


def ext_inv(d, a):
    a, out = np.array(a, dtype=np.uint64), np.zeros(d, dtype=np.uint64)
    lib().ext_inv_canonical(d, _p64(a), _p64(out))
    return [int(v) for v in out]


def verify(air_id, pub, proof):
    """0 = accepted"""
    pub = np.ascontiguousarray(pub, dtype=np.uint64)
    buf = np.frombuffer(proof, dtype=np.uint8)
    return lib().stark_verify(air_id, _p64(pub), pub.size, _p8(buf), buf.size)


def blake3(data):
    out = np.zeros(32, dtype=np.uint8)
    lib().blake3_256(bytes(data), len(data), _p8(out))
    return out.tobytes()


def sha3(data):
    out = np.zeros(32, dtype=np.uint8)
    lib().sha3_256(bytes(data), len(data), _p8(out))
    return out.tobytes()


def ntt(a_mont, inverse=False):
    a = np.array(a_mont, dtype=np.uint64)
    lib().ntt_natural(_p64(a), a.size, int(inverse))
    return a


def lde_column(col_mont, blowup):
    col = np.ascontiguousarray(col_mont, dtype=np.uint64)
    out = np.zeros(col.size * blowup, dtype=np.uint64)
    lib().lde_column(_p64(col), col.size, blowup, _p64(out))
    return out


def hash_elements(elems_mont, hash_fn=HASH_BLAKE3_256):
    e = np.ascontiguousarray(elems_mont, dtype=np.uint64)
    out = np.zeros(32, dtype=np.uint8)
    lib().hash_elements(hash_fn, _p64(e), e.size, _p8(out))
    return out.tobytes()


def merkle_nodes(leaves, hash_fn=HASH_BLAKE3_256):
    lv = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
    nodes = np.zeros((2 * lv.shape[0], 32), dtype=np.uint8)
    lib().merkle_build(hash_fn, _p8(lv), lv.shape[0], _p8(nodes))
    return nodes


def fri_fold4(evals_mont, alpha_mont):
    e = np.ascontiguousarray(evals_mont, dtype=np.uint64)
    out = np.zeros(e.size // 4, dtype=np.uint64)
    lib().fri_fold4(_p64(e), e.size, int(alpha_mont), _p64(out))
    return out


def rescue_permutation(state_canon):
    s = to_mont(state_canon)
    lib().rescue_apply_permutation(_p64(s))
    return from_mont(s)


def rescue_round(state_mont, step):
    s = np.array(state_mont, dtype=np.uint64)
    lib().rescue_apply_round(_p64(s), step)
    return s


def rescue_merge(a_canon, b_canon):
    a, b, o = to_mont(a_canon), to_mont(b_canon), np.zeros(7, dtype=np.uint64)
    lib().rescue_merge(_p64(a), _p64(b), _p64(o))
    return from_mont(o)


def rescue_digest(data_canon):
    d, o = to_mont(data_canon), np.zeros(7, dtype=np.uint64)
    lib().rescue_digest(_p64(d), d.size, _p64(o))
    return from_mont(o)


class Air:
    def __init__(self, air_id, trace_len, pub):
        pub = np.ascontiguousarray(pub, dtype=np.uint64)
        self.h = lib().air_new(air_id, trace_len, _p64(pub), pub.size)
        if not self.h:
            raise RuntimeError("air_new failed")
        self.air_id = air_id

    def __del__(self):
        if getattr(self, "h", None):
            lib().air_free(self.h)
            self.h = None

    def ce_blowup(self):
        return lib().air_ce_blowup(self.h)

    def eval_row(self, step, cur_mont, next_mont, num_constraints):
        cur = np.ascontiguousarray(cur_mont, dtype=np.uint64)
        nxt = np.ascontiguousarray(next_mont, dtype=np.uint64)
        res = np.zeros(num_constraints, dtype=np.uint64)
        lib().air_eval_row(self.h, step, _p64(cur), _p64(nxt), _p64(res))
        return res


def check_trace(air_id, trace_canon, pub, num_constraints):
    """Evaluate every transition constraint on every row pair of a trace (winterfell's debug `trace.validate`).
    Returns the list of (step, constraint) that do not vanish (empty = valid trace)."""
    tm = to_mont(trace_canon)
    n = tm.shape[1]
    air = Air(air_id, n, pub)
    bad = []
    cols = np.ascontiguousarray(tm.T)
    for s in range(n - 1):
        r = air.eval_row(s, cols[s], cols[s + 1], num_constraints)
        nz = np.nonzero(r)[0]
        bad.extend((s, int(i)) for i in nz)
    return bad
