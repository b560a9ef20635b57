"""Regenerates tests/golden/ref_vectors.npz from the REFERENCE's own functions.

Run in the build container (needs /root/reference and an AVX-512 host):
    make -C oracle && python tests/golden/make_golden.py
Every array named ref_* is an output of code compiled from /root/reference/{npj,cpra2,phj}.cpp
(through oracle/_ref/libref_*.so, see oracle/ref_shim/); in_* arrays are the inputs they were
run on.  tests/test_oracle_golden.py replays the oracle on the same inputs wherever the repo
goes (the GPU box has no /root/reference)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _oracle import (_p, aligned_u32, oracle_generate, ref, ref_available, sort_rows)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK = 65536


def al(a):
    b = aligned_u32(a.size)
    b[:] = a
    return b


def main():
    assert ref_available(), "oracle/_ref missing or host lacks avx512f"
    npj, cpra, phj = ref("npj"), ref("cpra"), ref("phj")
    for f in (npj.hjref_npj_join, cpra.hjref_cpra_join, phj.hjref_phj_local_join):
        f.restype = C.c_size_t
    g = {}
    # --- MT19937 streams (npj.cpp:138-175)
    seeds = np.array([1, 123, 0xDEADBEEF], np.uint32)
    streams = np.empty((3, 1500), np.uint32)
    for i, s in enumerate(seeds):
        row = np.empty(1500, np.uint32)
        npj.hjref_rand32_stream(C.c_uint32(int(s)), _p(row), C.c_size_t(1500))
        streams[i] = row
    g["rand32_seeds"], g["ref_rand32_streams"] = seeds, streams
    # --- shuffle / unique (npj.cpp:558-600)
    data = np.arange(1, 1001, dtype=np.uint32)
    npj.hjref_shuffle(_p(data), C.c_size_t(1000), C.c_uint32(77))
    g["ref_shuffle_1000_seed77"] = data
    buckets = 2003  # prime
    table = np.zeros(buckets, np.uint32)
    uniq = np.empty(1000, np.uint32)
    npj.hjref_unique(_p(uniq), C.c_size_t(1000), _p(table), C.c_size_t(buckets),
                     C.c_uint32(0x9E3779B1), C.c_uint32(99))
    g["ref_unique_1000_seed99"] = uniq
    # --- small case: full inputs and full outputs
    nr, ns = 3000, 10000
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=3, seed=11)
    # CPRA first, on duplicate-free R: the reference's vector build does not survive duplicate
    # keys inside its tiny 4096-way partitions at this size (it crashes; observed, not analysed)
    bl = int(ns * 1.5 / BLOCK) + 3
    ko, so, ro = (aligned_u32(bl * BLOCK, 0) for _ in range(3))
    g["in_small_rk_nodup"], g["in_small_rv_nodup"] = rk.copy(), rv.copy()
    ark, arv, ask, asv = al(rk), al(rv), al(sk), al(sv)
    cnt = cpra.hjref_cpra_join(_p(ark), _p(arv), C.c_size_t(nr), _p(ask), _p(asv), C.c_size_t(ns),
                               C.c_int(5), _p(ko), _p(so), _p(ro), C.c_size_t(bl))
    g["ref_small_cpra_rows"] = sort_rows(ko[:cnt].copy(), so[:cnt].copy(), ro[:cnt].copy())
    # give R duplicates so that "emit every matching pair" is pinned too (NPJ, PHJ kernels)
    rk[:200] = rk[200:400]
    rv[:200] = rk[:200] * np.uint32(0x01000193)
    g["in_small_rk"], g["in_small_rv"], g["in_small_sk"], g["in_small_sv"] = rk, rv, sk, sv
    ark, arv, ask, asv = al(rk), al(rv), al(sk), al(sv)
    factor = 0x9E3779B1
    nb = int(nr / 0.9)
    tab = np.zeros(nb, np.uint64)
    npj.hjref_npj_build(_p(ark), _p(arv), C.c_size_t(nr), tab.ctypes.data_as(C.POINTER(C.c_uint64)),
                        C.c_size_t(nb), C.c_uint32(factor))
    g["small_npj_factor"], g["ref_small_npj_table"] = np.uint32(factor), tab
    ko[:] = 0; so[:] = 0; ro[:] = 0
    cnt = npj.hjref_npj_join(_p(ark), _p(arv), C.c_size_t(nr), _p(ask), _p(asv), C.c_size_t(ns),
                             C.c_size_t(nb), C.c_uint32(factor), _p(ko), _p(so), _p(ro), C.c_size_t(bl))
    g["ref_small_npj_rows"] = sort_rows(ko[:cnt].copy(), so[:cnt].copy(), ro[:cnt].copy())
    ark, arv, ask, asv = al(rk), al(rv), al(sk), al(sv)  # run_hj partitions in place of its inputs
    ko[:] = 0; so[:] = 0; ro[:] = 0
    jf = np.array([0x9E3779B1, 0x85EBCA6B], np.uint32)
    cnt = phj.hjref_phj_local_join(_p(ark), _p(arv), C.c_size_t(nr), _p(ask), _p(asv), C.c_size_t(ns),
                                   C.c_size_t(37), C.c_uint32(0x2545F491), _p(jf), _p(ko), _p(so),
                                   _p(ro), C.c_size_t(bl))
    g["ref_small_phj_rows"] = sort_rows(ko[:cnt].copy(), so[:cnt].copy(), ro[:cnt].copy())
    # --- histogram / partition (cpra2.cpp:801-1075, AVX-512 forms) on the small S column
    pf = 0x12345679
    g["part_factor"] = np.uint32(pf)
    for P in (64, 100, 4096):
        c = np.zeros(P, np.uint32)
        cpra.hjref_histogram(_p(ask), C.c_size_t(ns), _p(c), C.c_uint32(pf), C.c_size_t(P))
        g[f"ref_small_hist_{P}"] = c
        k1, v1 = aligned_u32(ns), aligned_u32(ns)
        cpra.hjref_partition(_p(ask), _p(asv), C.c_size_t(ns), _p(c), _p(k1), _p(v1),
                             C.c_uint32(pf), C.c_size_t(P))
        # order inside a partition is lane-dependent in the vector code: keep it canonical
        off = np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
        kk, vv = k1.copy(), v1.copy()
        for p in range(P):
            s = slice(off[p], off[p + 1])
            o = np.lexsort((vv[s], kk[s]))
            kk[s], vv[s] = kk[s][o], vv[s][o]
        g[f"ref_small_part_keys_{P}"], g[f"ref_small_part_vals_{P}"] = kk, vv
    # --- double-hash table (cpra2.cpp:307-398) for one partition-sized input
    nb2 = int(nr / 0.4) | 1
    while not cpra.hjref_odd_prime(C.c_uint64(nb2)):
        nb2 += 2
    t2 = np.zeros(nb2, np.uint64)
    urk = np.unique(rk)  # vector build resolves intra-vector collisions in lane order: use unique keys
    urv = urk * np.uint32(0x01000193)
    g["in_dh_keys"], g["in_dh_vals"] = urk, urv
    a1, a2 = al(urk), al(urv)
    cpra.hjref_dh_build_s(_p(a1), _p(a2), C.c_size_t(urk.size), t2.ctypes.data_as(C.POINTER(C.c_uint64)),
                          C.c_size_t(nb2), _p(jf))
    g["dh_factors"], g["ref_dh_table_scalar"] = jf, t2
    # --- medium cases: generator seed + the reference's count and checksums only
    med = []
    for (mr, ms, seed, T) in ((1 << 16, 1 << 18, 7, 2), (1 << 18, 1 << 18, 9, 4)):
        rk, rv, sk, sv, _, _ = oracle_generate(mr, ms, threads=T, seed=seed)
        ark, arv, ask, asv = al(rk), al(rv), al(sk), al(sv)
        bl = int(ms * 1.05 / BLOCK) + 3
        ko, so, ro = (aligned_u32(bl * BLOCK, 0) for _ in range(3))
        cnt = npj.hjref_npj_join(_p(ark), _p(arv), C.c_size_t(mr), _p(ask), _p(asv), C.c_size_t(ms),
                                 C.c_size_t(int(mr / 0.9)), C.c_uint32(factor), _p(ko), _p(so), _p(ro),
                                 C.c_size_t(bl))
        sums = [int(x[:cnt].astype(np.uint64).sum(dtype=np.uint64)) for x in (ko, so, ro)]
        ko[:] = 0; so[:] = 0; ro[:] = 0
        cnt2 = cpra.hjref_cpra_join(_p(ark), _p(arv), C.c_size_t(mr), _p(ask), _p(asv), C.c_size_t(ms),
                                    C.c_int(3), _p(ko), _p(so), _p(ro), C.c_size_t(bl))
        sums2 = [int(x[:cnt2].astype(np.uint64).sum(dtype=np.uint64)) for x in (ko, so, ro)]
        assert (cnt, sums) == (cnt2, sums2), "reference NPJ and CPRA disagree"
        med.append([mr, ms, seed, T, cnt] + sums)
    g["ref_medium"] = np.array(med, dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **g)
    print("wrote", os.path.join(HERE, "ref_vectors.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
