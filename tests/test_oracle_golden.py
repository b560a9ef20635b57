"""Pins oracle/hj_oracle.c against tests/golden/ref_vectors.npz -- outputs of the REFERENCE's
own functions (compiled from /root/reference by oracle/Makefile, captured by
tests/golden/make_golden.py).  Runs anywhere: needs neither /root/reference nor a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from _oracle import _p, lib, numpy_join, oracle_generate, oracle_join, sort_rows

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))


class R32(C.Structure):
    _fields_ = [("num", C.c_uint32 * 625), ("index", C.c_size_t)]


def _stream(seed, n):
    L = lib()
    L.hjo_rand32_next.restype = C.c_uint32
    st = R32()
    L.hjo_rand32_seed(C.byref(st), C.c_uint32(seed))
    return st, np.array([L.hjo_rand32_next(C.byref(st)) for _ in range(n)], np.uint32)


def test_mt19937_matches_reference():
    for seed, want in zip(G["rand32_seeds"], G["ref_rand32_streams"]):
        _, got = _stream(int(seed), want.size)
        assert (got == want).all()


def test_shuffle_and_unique_match_reference():
    L = lib()
    st = R32()
    L.hjo_rand32_seed(C.byref(st), C.c_uint32(77))
    data = np.arange(1, 1001, dtype=np.uint32)
    L.hjo_shuffle(_p(data), C.c_size_t(1000), C.byref(st))
    assert (data == G["ref_shuffle_1000_seed77"]).all()
    L.hjo_rand32_seed(C.byref(st), C.c_uint32(99))
    table = np.zeros(2003, np.uint32)
    uniq = np.empty(1000, np.uint32)
    L.hjo_unique(_p(uniq), C.c_size_t(1000), _p(table), C.c_size_t(2003), C.c_uint32(0x9E3779B1),
                 C.c_uint32(0), C.byref(st))
    assert (uniq == G["ref_unique_1000_seed99"]).all()


def test_npj_table_bit_identical_to_reference_build():
    rk, rv = G["in_small_rk"], G["in_small_rv"]
    want = G["ref_small_npj_table"]
    tab = np.zeros(want.size, np.uint64)
    lib().hjo_npj_build(_p(rk), _p(rv), rk.size, tab.ctypes.data_as(C.POINTER(C.c_uint64)),
                        want.size, int(G["small_npj_factor"]), 0)
    assert (tab == want).all()


@pytest.mark.parametrize("algo,rows,rkey", [("npj", "ref_small_npj_rows", "in_small_rk"),
                                            ("phj", "ref_small_phj_rows", "in_small_rk"),
                                            ("cpra", "ref_small_npj_rows", "in_small_rk"),
                                            ("cpra", "ref_small_cpra_rows", "in_small_rk_nodup"),
                                            ("npj", "ref_small_cpra_rows", "in_small_rk_nodup")])
@pytest.mark.parametrize("threads", [1, 3])
def test_join_rows_match_reference(algo, rows, rkey, threads):
    rk = G[rkey]
    rv = G[rkey.replace("rk", "rv")]
    got = oracle_join(algo, rk, rv, G["in_small_sk"], G["in_small_sv"], threads=threads)
    want = G[rows]
    assert got.count == want.shape[0]
    assert (got.sorted_rows() == want).all()
    assert got.checks() == numpy_join(rk, rv, G["in_small_sk"], G["in_small_sv"]).checks()


@pytest.mark.parametrize("P", [64, 100, 4096])
def test_histogram_and_partition_match_reference(P):
    L = lib()
    sk, sv = G["in_small_sk"], G["in_small_sv"]
    f = int(G["part_factor"])
    c = np.zeros(P, np.uint32)
    L.hjo_histogram(_p(sk), sk.size, _p(c), f, P)
    assert (c == G[f"ref_small_hist_{P}"]).all()
    k, v = np.empty_like(sk), np.empty_like(sv)
    L.hjo_partition(_p(sk), _p(sv), sk.size, _p(c), _p(k), _p(v), f, P)
    off = np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
    for p in range(P):  # the vector reference's order inside a partition is lane-dependent
        s = slice(off[p], off[p + 1])
        o = np.lexsort((v[s], k[s]))
        k[s], v[s] = k[s][o], v[s][o]
    assert (k == G[f"ref_small_part_keys_{P}"]).all()
    assert (v == G[f"ref_small_part_vals_{P}"]).all()


def test_double_hash_table_bit_identical_to_reference_build_s():
    want = G["ref_dh_table_scalar"]
    keys, vals = G["in_dh_keys"], G["in_dh_vals"]
    tab = np.zeros(want.size, np.uint64)
    lib().hjo_dh_build(_p(keys), _p(vals), keys.size, tab.ctypes.data_as(C.POINTER(C.c_uint64)),
                       want.size, _p(G["dh_factors"]), 0)
    assert (tab == want).all()


@pytest.mark.parametrize("row", range(2))
@pytest.mark.parametrize("algo", ["npj", "phj", "cpra"])
def test_medium_checksums_match_reference(row, algo):
    nr, ns, seed, T, cnt, s_key, s_outer, s_inner = (int(x) for x in G["ref_medium"][row])
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=T, seed=seed)
    got = oracle_join(algo, rk, rv, sk, sv, threads=4, materialize=False)
    assert got.checks() == (cnt, s_key, s_outer, s_inner)


def test_planner_matches_reference_constants():
    fan = (C.c_size_t * 6)()
    assert lib().hjo_plan_fanout(4096, fan) == 2 and list(fan[:3]) == [64, 64, 1]  # cpra2.cpp:2023
    assert lib().hjo_plan_fanout(327, fan) == 1 and list(fan[:2]) == [327, 1]
    assert lib().hjo_plan_fanout(5, fan) == 0 and fan[0] == 1


def test_oracle_rejects_empty_sentinel_key():
    k = np.array([0, 5], np.uint32)
    with pytest.raises(ValueError):
        oracle_join("npj", k, k, k, k)


def test_sorted_rows_helper():
    k = np.array([3, 1, 3], np.uint32)
    assert sort_rows(k, k, k)[0, 0] == 1
