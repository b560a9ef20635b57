"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the reference's
own functions compiled into oracle/_ref/.  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

u32p = C.POINTER(C.c_uint32)


def _p(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


def aligned_u32(n, fill=None):
    """64-byte aligned uint32 array (the reference asserts alignment, npj.cpp:226-227)."""
    raw = np.empty(n * 4 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + n * 4].view(np.uint32)
    if fill is not None:
        a[:] = fill
    return a


class HjoResult(C.Structure):
    _fields_ = [("count", C.c_uint64), ("sum_key", C.c_uint64), ("sum_outer", C.c_uint64),
                ("sum_inner", C.c_uint64), ("keys", u32p), ("outer_vals", u32p),
                ("inner_vals", u32p), ("seconds", C.c_double)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "hj_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        for name in ("hjo_npj", "hjo_phj", "hjo_cpra"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [u32p, u32p, C.c_size_t, u32p, u32p, C.c_size_t, C.c_int, C.c_uint32,
                          C.c_int, C.POINTER(HjoResult)]
        L.hjo_result_free.argtypes = [C.POINTER(HjoResult)]
        L.hjo_hash.restype = C.c_uint32
        L.hjo_hash.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.hjo_generate.restype = C.c_int
        L.hjo_generate.argtypes = [C.c_size_t, C.c_size_t, C.c_double, C.c_int, C.c_uint32,
                                   u32p, u32p, u32p, u32p, u32p, u32p]
        L.hjo_histogram.argtypes = [u32p, C.c_size_t, u32p, C.c_uint32, C.c_size_t]
        L.hjo_partition.argtypes = [u32p, u32p, C.c_size_t, u32p, u32p, u32p, C.c_uint32, C.c_size_t]
        L.hjo_plan_fanout.restype = C.c_size_t
        L.hjo_plan_fanout.argtypes = [C.c_size_t, C.POINTER(C.c_size_t)]
        L.hjo_odd_prime.restype = C.c_int
        L.hjo_odd_prime.argtypes = [C.c_uint64]
        L.hjo_npj_build.argtypes = [u32p, u32p, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t,
                                    C.c_uint32, C.c_uint32]
        L.hjo_dh_build.argtypes = [u32p, u32p, C.c_size_t, C.POINTER(C.c_uint64), C.c_size_t,
                                   u32p, C.c_uint32]
        L.hjo_relation_write.restype = C.c_int
        L.hjo_relation_write.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, u32p, u32p]
        L.hjo_relation_read.restype = C.c_int
        L.hjo_relation_read.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, u32p, u32p]
        _lib = L
    return _lib


class JoinResult:
    def __init__(self, count, sum_key, sum_outer, sum_inner, rows=None, seconds=0.0):
        self.count, self.sum_key, self.sum_outer, self.sum_inner = count, sum_key, sum_outer, sum_inner
        self.rows = rows  # (keys, outer_vals, inner_vals) or None
        self.seconds = seconds

    def checks(self):
        return (self.count, self.sum_key, self.sum_outer, self.sum_inner)

    def sorted_rows(self):
        return sort_rows(*self.rows)


def sort_rows(k, o, i):
    """rows sorted lexicographically by (key, outer_val, inner_val)."""
    order = np.lexsort((i, o, k))
    return np.stack([k[order], o[order], i[order]], axis=1)


def oracle_join(algo, rk, rv, sk, sv, threads=1, seed=1, materialize=True):
    L = lib()
    res = HjoResult()
    rk, rv, sk, sv = (np.ascontiguousarray(x, dtype=np.uint32) for x in (rk, rv, sk, sv))
    rc = getattr(L, "hjo_" + algo)(_p(rk), _p(rv), rk.size, _p(sk), _p(sv), sk.size,
                                   threads, seed, int(materialize), C.byref(res))
    if rc != 0:
        raise ValueError(f"oracle {algo} rejected the input (rc={rc})")
    rows = None
    if materialize:
        n = res.count
        rows = tuple(np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.empty(0, np.uint32)
                     for p in (res.keys, res.outer_vals, res.inner_vals))
    out = JoinResult(res.count, res.sum_key, res.sum_outer, res.sum_inner, rows, res.seconds)
    L.hjo_result_free(C.byref(res))
    return out


def oracle_generate(nr, ns, selectivity=1.0, threads=1, seed=42):
    L = lib()
    rk, rv, sk, sv = (np.empty(n, np.uint32) for n in (nr, nr, ns, ns))
    fi, fo = C.c_uint32(), C.c_uint32()
    rc = L.hjo_generate(nr, ns, selectivity, threads, seed, _p(rk), _p(rv), _p(sk), _p(sv),
                        C.byref(fi), C.byref(fo))
    assert rc == 0
    return rk, rv, sk, sv, fi.value, fo.value


def numpy_join(rk, rv, sk, sv, materialize=True):
    """Independent exact equi-join (sort + searchsorted), all (r,s) pairs with equal keys."""
    rk, rv, sk, sv = (np.asarray(x, dtype=np.uint32) for x in (rk, rv, sk, sv))
    order = np.argsort(rk, kind="stable")
    rks, rvs = rk[order], rv[order]
    lo = np.searchsorted(rks, sk, side="left")
    hi = np.searchsorted(rks, sk, side="right")
    reps = (hi - lo).astype(np.int64)
    count = int(reps.sum())
    s_idx = np.repeat(np.arange(sk.size), reps)
    starts = np.repeat(lo, reps)
    within = np.arange(count) - np.repeat(np.cumsum(reps) - reps, reps)
    r_idx = starts + within
    k, o, i = sk[s_idx], sv[s_idx], rvs[r_idx]
    sums = tuple(int(x.astype(np.uint64).sum(dtype=np.uint64)) for x in (k, o, i))
    return JoinResult(count, *sums, rows=(k, o, i) if materialize else None)


# ----------------------------------------------------------------------------- oracle/_ref

def ref_available():
    if not os.path.exists(os.path.join(REF_DIR, "libref_npj.so")):
        return False
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


_ref = {}


def ref(name):
    if name not in _ref:
        _ref[name] = C.CDLL(os.path.join(REF_DIR, f"libref_{name}.so"))
    return _ref[name]
