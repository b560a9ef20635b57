"""Live pin: oracle/hj_oracle.c against the reference's own compiled functions in oracle/_ref/
(only where oracle/_ref was built -- i.e. where /root/reference exists -- and the host has
AVX-512F).  tests/test_oracle_golden.py holds the same comparisons as committed fixtures."""
import ctypes as C

import numpy as np
import pytest

from _oracle import (_p, aligned_u32, oracle_generate, oracle_join, ref, ref_available, sort_rows)

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built or no avx512f")


def _al(a):
    b = aligned_u32(a.size)
    b[:] = a
    return b


@pytest.mark.parametrize("nr,ns,seed", [(1 << 14, 1 << 16, 3), (50000, 50000, 4), (1 << 16, 1 << 14, 5)])
def test_reference_npj_cpra_phj_agree_with_oracle(nr, ns, seed, capfd):
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=2, seed=seed)
    want = oracle_join("npj", rk, rv, sk, sv, threads=2)
    bl = int(want.count * 1.05 / 65536) + 3
    ko, so, ro = (aligned_u32(bl * 65536, 0) for _ in range(3))
    npj, cpra, phj = ref("npj"), ref("cpra"), ref("phj")
    for f in (npj.hjref_npj_join, cpra.hjref_cpra_join, phj.hjref_phj_local_join):
        f.restype = C.c_size_t
    ark, arv, ask, asv = _al(rk), _al(rv), _al(sk), _al(sv)
    cnt = npj.hjref_npj_join(_p(ark), _p(arv), C.c_size_t(nr), _p(ask), _p(asv), C.c_size_t(ns),
                             C.c_size_t(int(nr / 0.9)), C.c_uint32(0x9E3779B1), _p(ko), _p(so), _p(ro),
                             C.c_size_t(bl))
    assert cnt == want.count
    assert (sort_rows(ko[:cnt], so[:cnt], ro[:cnt]) == want.sorted_rows()).all()
    ko[:] = 0
    jf = np.array([0x9E3779B1, 0x85EBCA6B], np.uint32)
    cnt = phj.hjref_phj_local_join(_p(ark), _p(arv), C.c_size_t(nr), _p(ask), _p(asv), C.c_size_t(ns),
                                   C.c_size_t(53), C.c_uint32(0x2545F491), _p(jf), _p(ko), _p(so),
                                   _p(ro), C.c_size_t(bl))
    assert cnt == want.count
    assert (sort_rows(ko[:cnt], so[:cnt], ro[:cnt]) == want.sorted_rows()).all()
    ko[:] = 0
    cnt = cpra.hjref_cpra_join(_p(ark), _p(arv), C.c_size_t(nr), _p(ask), _p(asv), C.c_size_t(ns),
                               C.c_int(seed), _p(ko), _p(so), _p(ro), C.c_size_t(bl))
    assert cnt == want.count
    assert (sort_rows(ko[:cnt], so[:cnt], ro[:cnt]) == want.sorted_rows()).all()
