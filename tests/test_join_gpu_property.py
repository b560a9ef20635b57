"""Property test: for ANY pair of small relations -- sizes from 0 to a few thousand, key domains
small enough to force equal keys on both sides, the reserved-looking values 0 and 0xFFFFFFFF
mixed in, any radix plan -- NPJ and PHJ through the C ABI emit exactly the rows of a numpy
equi-join (every (r, s) pair with r.key == s.key, npj.cpp:288-290), on the device and the host
entry points."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import hash_join_codes_knl_b200 as hj
from _oracle import numpy_join, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = hj.Engine(0)
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


PLANS = [{}, {"radix_bits": (3,)}, {"radix_bits": (8, 8)}, {"radix_bits": (5, 6, 5)}, {"radix_bits": (9, 2)}, {"part_tuples": 64}, {"seed": 7}]


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(nr=st.integers(0, 3000), ns=st.integers(0, 6000), domain_log2=st.integers(1, 32), special=st.booleans(),
       plan=st.sampled_from(PLANS), algo=st.sampled_from(["npj", "phj"]), host=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_any_small_join_equals_numpy(eng, nr, ns, domain_log2, special, plan, algo, host, seed):
    rng = np.random.default_rng(seed)
    hi = 1 << domain_log2
    rk = rng.integers(0, hi, nr, dtype=np.uint64).astype(np.uint32)
    sk = rng.integers(0, hi, ns, dtype=np.uint64).astype(np.uint32)
    if special:
        for arr in (rk, sk):
            if arr.size:
                arr[rng.integers(0, arr.size, max(1, arr.size // 50))] = rng.choice(np.array([0, 0xFFFFFFFF], np.uint32))
    rv = rng.integers(0, 1 << 32, nr, dtype=np.uint64).astype(np.uint32)
    sv = rng.integers(0, 1 << 32, ns, dtype=np.uint64).astype(np.uint32)
    if special and nr and ns:
        rv[0] = sv[0] = rk[0] = sk[0] = 0xFFFFFFFF          # the pair that looks like an empty table slot
    ur, cr = np.unique(rk, return_counts=True)
    us, cs = np.unique(sk, return_counts=True)
    _, ir, is_ = np.intersect1d(ur, us, return_indices=True)
    if int((cr[ir].astype(np.int64) * cs[is_]).sum()) > 3_000_000:   # tiny domains on both sides: quadratic blow-up, covered elsewhere
        return
    want = numpy_join(rk, rv, sk, sv)
    if host:
        got = getattr(eng, algo)((rk, rv), (sk, sv), **plan)
    else:
        got = getattr(eng, algo)((dev(rk), dev(rv)), (dev(sk), dev(sv)), **plan)
    assert got.checks() == want.checks()
    assert (sort_rows(*got.rows_numpy()) == want.sorted_rows()).all()
