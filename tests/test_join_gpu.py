"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle
(oracle/hj_oracle.c, pinned to the reference by tests/test_oracle_golden.py), the committed
golden rows of the reference itself, and an independent numpy join -- bit-exact: count, the
three uint64 checksums, and the sorted materialised rows."""
import os

import numpy as np
import pytest
import torch

import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
from _oracle import lib as olib, _p, numpy_join, oracle_generate, oracle_join, sort_rows

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))


@pytest.fixture(scope="module")
def eng():
    e = hj.Engine(0)
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def run(eng, algo, rk, rv, sk, sv, where="device", **opts):
    if where == "device":
        return getattr(eng, algo)((dev(rk), dev(rv)), (dev(sk), dev(sv)), **opts)
    return getattr(eng, algo)((rk, rv), (sk, sv), **opts)


def assert_same(got, want, rows=True):
    assert got.checks() == want.checks()
    if rows:
        assert (sort_rows(*got.rows_numpy()) == want.sorted_rows()).all()


ALGOS = ["npj", "phj"]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("nr,ns,seed", [(1000, 1000, 1), (3000, 10000, 2), (1 << 16, 1 << 18, 3),
                                        (100003, 70001, 4), (1 << 20, 1 << 20, 5), (1 << 18, 1 << 22, 6)])
def test_join_matches_oracle(eng, algo, nr, ns, seed):
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=3, seed=seed)
    want = oracle_join(algo, rk, rv, sk, sv, threads=4)
    assert_same(run(eng, algo, rk, rv, sk, sv), want)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("rkey,rows", [("in_small_rk", "ref_small_npj_rows"), ("in_small_rk", "ref_small_phj_rows"),
                                       ("in_small_rk_nodup", "ref_small_cpra_rows")])
def test_join_matches_reference_golden_rows(eng, algo, rkey, rows):
    """rows produced by the reference's own build/probe/run_hj (tests/golden/make_golden.py)"""
    got = run(eng, algo, G[rkey], G[rkey.replace("rk", "rv")], G["in_small_sk"], G["in_small_sv"])
    assert got.count == G[rows].shape[0]
    assert (sort_rows(*got.rows_numpy()) == G[rows]).all()


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("row", range(2))
def test_join_matches_reference_golden_checksums(eng, algo, row):
    nr, ns, seed, T, cnt, s_key, s_outer, s_inner = (int(x) for x in G["ref_medium"][row])
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=T, seed=seed)
    assert run(eng, algo, rk, rv, sk, sv).checks() == (cnt, s_key, s_outer, s_inner)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("nr,ns", [(0, 0), (0, 100), (100, 0), (1, 1), (5, 3), (31, 33), (17, 1025), (4097, 15)])
def test_tiny_and_ragged_sizes(eng, algo, nr, ns):
    rng = np.random.default_rng(nr * 1000 + ns)
    rk = rng.integers(1, 50, nr, dtype=np.uint32)          # many duplicates on both sides
    sk = rng.integers(1, 50, ns, dtype=np.uint32)
    rv, sv = rk * np.uint32(3) + np.uint32(1), sk * np.uint32(5) + np.uint32(2)
    want = numpy_join(rk, rv, sk, sv)
    for where in ("device", "host"):
        assert_same(run(eng, algo, rk, rv, sk, sv, where=where), want)


# {} lets the planner choose (small inputs: HASH tables); 16 radix bits leave 16 hash bits, which
# selects the DIRECT (bitmap + rank) tables of csrc/part_join.cu
PLANS = [{}, {"radix_bits": (8, 8)}, {"radix_bits": (6, 5, 5)}]


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("algo", ALGOS)
def test_special_key_values(eng, algo, plan):
    """key 0 (the reference's empty sentinel, npj.cpp:205), 0xFFFFFFFF, and the one pair
    (0xFFFFFFFF, 0xFFFFFFFF) that looks like an empty slot of this engine's tables"""
    rk = np.array([0, 0, 1, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, 7, 7], np.uint32)
    rv = np.array([0, 9, 1, 0xFFFFFFFF, 0xFFFFFFFF, 5, 0xFFFFFFFF, 3, 0, 0xFFFFFFFF], np.uint32)
    sk = np.array([0, 0xFFFFFFFF, 7, 2, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE, 0], np.uint32)
    sv = np.array([11, 0xFFFFFFFF, 13, 14, 15, 16, 0xFFFFFFFF, 0], np.uint32)
    want = numpy_join(rk, rv, sk, sv)
    assert want.count == 2 * 2 + 3 * 2 + 2 + 1 + 1      # keys 0, 0xFFFFFFFF, 7, 0x80000000, 0xFFFFFFFE
    assert_same(run(eng, algo, rk, rv, sk, sv, **plan), want)
    # and at scale: many sentinel pairs on the build side
    rng = np.random.default_rng(5)
    rk = np.concatenate([np.full(300, 0xFFFFFFFF, np.uint32), rng.integers(0, 1 << 32, 5000, dtype=np.uint32)])
    rv = np.concatenate([np.full(300, 0xFFFFFFFF, np.uint32), rng.integers(0, 1 << 32, 5000, dtype=np.uint32)])
    sk = np.concatenate([np.full(40, 0xFFFFFFFF, np.uint32), rk[300:2000], rng.integers(0, 1 << 32, 3000, dtype=np.uint32)])
    sv = np.arange(sk.size, dtype=np.uint32)
    assert_same(run(eng, algo, rk, rv, sk, sv, **plan), numpy_join(rk, rv, sk, sv))


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("algo", ALGOS)
def test_duplicate_heavy_build_side_overflows_capacity(eng, algo, plan):
    """every pair is emitted (no _UNIQUE, npj.cpp:288-290): 3000 x 2000 equal keys -> 6M rows from
    5000 input tuples, which overflows the shared-memory stage and the default result capacity"""
    rk = np.full(3000, 77, np.uint32)
    rv = np.arange(3000, dtype=np.uint32)
    sk = np.concatenate([np.full(2000, 77, np.uint32), np.arange(100, 600, dtype=np.uint32)])
    sv = np.arange(sk.size, dtype=np.uint32) * np.uint32(7)
    want = numpy_join(rk, rv, sk, sv)
    assert want.count == 6_000_000
    assert_same(run(eng, algo, rk, rv, sk, sv, **plan), want)


@pytest.mark.parametrize("plan", PLANS)
def test_some_partitions_with_equal_build_keys(eng, plan):
    """a large build side in which a few keys repeat: most partition fills take the DIRECT tables,
    the fills holding an equal pair fall back to HASH tables inside the same kernel"""
    rk, rv, sk, sv, _, _ = oracle_generate(1 << 19, 1 << 20, threads=2, seed=17)
    rk[1000:1400] = rk[5000:5400]                     # 400 keys now appear twice
    rk[2000:2010] = rk[7000]                          # one key eleven times
    rv = np.arange(rk.size, dtype=np.uint32)          # payloads no longer a function of the key
    want = numpy_join(rk, rv, sk, sv)
    assert_same(run(eng, "phj", rk, rv, sk, sv, **plan), want)
    assert_same(run(eng, "npj", rk, rv, sk, sv), want)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("sel", [0.0, 0.5, 1.0])
def test_selectivity(eng, algo, sel):
    rk, rv, sk, sv, _, _ = oracle_generate(1 << 15, 1 << 15, selectivity=sel, threads=2, seed=8)
    want = oracle_join(algo, rk, rv, sk, sv, threads=2)
    assert want.count == int((1 << 15) * sel)
    assert_same(run(eng, algo, rk, rv, sk, sv), want)


@pytest.mark.parametrize("algo", ALGOS)
def test_heavy_hitter_probe_keys(eng, algo):
    """one probe key holds half of S: its partition is cut into many tasks (csrc/part_join.cu)"""
    rk, rv, sk, sv, _, _ = oracle_generate(1 << 14, 1 << 20, threads=2, seed=9)
    sk[::2] = rk[5]
    sv[::2] = np.arange(sk.size // 2, dtype=np.uint32)
    want = numpy_join(rk, rv, sk, sv, materialize=False)
    got = run(eng, algo, rk, rv, sk, sv)
    assert got.checks() == want.checks()
    got2 = run(eng, algo, rk, rv, sk, sv, materialize=False)
    assert got2.checks() == want.checks() and not got2.materialized


@pytest.mark.parametrize("algo", ALGOS)
def test_host_entry_equals_device_entry(eng, algo):
    rk, rv, sk, sv, _, _ = oracle_generate(50000, 200000, threads=2, seed=10)
    a = run(eng, algo, rk, rv, sk, sv, where="device")
    b = run(eng, algo, rk, rv, sk, sv, where="host")
    assert a.checks() == b.checks() and b.seconds_e2e > 0 and not b.rows_on_device
    assert (sort_rows(*a.rows_numpy()) == sort_rows(*b.rows_numpy())).all()


@pytest.fixture
def small_host_slices(monkeypatch):
    """the host entry points slice the probe side from 2 x HJB_HOST_SLICE tuples on (default 2^22)"""
    monkeypatch.setenv("HJB_HOST_SLICE", "4096")


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("nr,ns", [(20000, 8192), (50000, 200001), (3000, 70000), (1500, 70001), (1 << 17, 1 << 20)])
def test_pipelined_host_entry_matches_oracle(eng, small_host_slices, algo, nr, ns):
    """probe side copied, joined and copied back slice by slice on three streams: same rows"""
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=2, seed=21)
    sk[::5] ^= np.uint32(0x40000000)                         # some probe keys without a partner
    want = oracle_join(algo, rk, rv, sk, sv, threads=2)
    got = run(eng, algo, rk, rv, sk, sv, where="host")
    assert not got.rows_on_device and got.seconds_e2e > 0
    assert_same(got, want)
    nomat = run(eng, algo, rk, rv, sk, sv, where="host", materialize=False)
    assert nomat.checks() == want.checks() and not nomat.materialized


@pytest.mark.parametrize("algo", ALGOS)
def test_pipelined_host_entry_falls_back_when_rows_exceed_capacity(eng, small_host_slices, algo):
    """equal build keys: more rows than max(|R|, |S|) -> the sliced path hands over to the plain one"""
    rk = np.concatenate([np.full(40, 5, np.uint32), np.arange(100, 1100, dtype=np.uint32)])
    rv = np.arange(rk.size, dtype=np.uint32)
    sk = np.concatenate([np.full(3000, 5, np.uint32), np.arange(0, 30000, dtype=np.uint32)])
    sv = np.arange(sk.size, dtype=np.uint32) * np.uint32(3)
    want = numpy_join(rk, rv, sk, sv)
    assert want.count > sk.size
    assert_same(run(eng, algo, rk, rv, sk, sv, where="host"), want)


@pytest.mark.parametrize("algo", ALGOS)
def test_pipelined_and_plain_host_entries_agree_at_default_slicing(eng, algo):
    """2^23 probe tuples: two slices of 2^22 without any override"""
    rk, rv, sk, sv, _, _ = oracle_generate(1 << 20, 1 << 23, threads=4, seed=22)
    want = oracle_join(algo, rk, rv, sk, sv, threads=4, materialize=False)
    got = run(eng, algo, rk, rv, sk, sv, where="host")
    assert got.checks() == want.checks()
    k, o, i = got.rows_numpy()
    assert k.size == want.count and int(k.astype(np.uint64).sum()) == want.sum_key
    assert int(o.astype(np.uint64).sum()) == want.sum_outer and int(i.astype(np.uint64).sum()) == want.sum_inner


def test_repeated_phj_on_the_same_buffers_replays_a_graph_and_rereads_the_data(eng):
    """second call on identical buffers captures the launch sequence, later calls replay it; the
    replay must see new CONTENTS of those buffers, and a different size must leave the graph behind"""
    rk, rv, sk, sv, _, _ = oracle_generate(1 << 16, 1 << 18, threads=2, seed=31)
    drk, drv, dsk, dsv = dev(rk), dev(rv), dev(sk), dev(sv)
    want = oracle_join("phj", rk, rv, sk, sv, threads=2)
    for _ in range(4):
        assert_same(eng.phj((drk, drv), (dsk, dsv)), want)
    sk2 = sk.copy()
    sk2[::3] ^= np.uint32(0x20000000)
    dsk.copy_(dev(sk2))                                         # same pointer, new keys
    want2 = oracle_join("phj", rk, rv, sk2, sv, threads=2)
    assert want2.count != want.count
    for _ in range(2):
        assert_same(eng.phj((drk, drv), (dsk, dsv)), want2)
    assert_same(eng.phj((drk, drv), (dsk[:100000], dsv[:100000])), oracle_join("phj", rk, rv, sk2[:100000], sv[:100000], threads=2))
    assert_same(eng.phj((drk, drv), (dsk, dsv)), want2)
    # a build side so small that the plan has no radix pass (one partition, offsets set by a tiny kernel)
    rk4, rv4, sk4, sv4, _, _ = oracle_generate(1000, 50000, threads=2, seed=32)
    d4 = dev(rk4), dev(rv4)
    s4 = dev(sk4), dev(sv4)
    want4 = oracle_join("phj", rk4, rv4, sk4, sv4, threads=2)
    for _ in range(3):
        assert_same(eng.phj(d4, s4), want4)
    # equal build keys: the replayed graph overflows the result capacity -> eager rerun with a larger buffer
    rk3 = np.full(3000, 9, np.uint32)
    d3 = dev(rk3), dev(np.arange(3000, dtype=np.uint32))
    s3k = np.concatenate([np.full(1500, 9, np.uint32), np.arange(100, 4000, dtype=np.uint32)])
    s3 = dev(s3k), dev(np.arange(s3k.size, dtype=np.uint32))
    want3 = numpy_join(rk3, np.arange(3000, dtype=np.uint32), s3k, np.arange(s3k.size, dtype=np.uint32))
    for _ in range(3):
        got = eng.phj(d3, s3)
        assert got.checks() == want3.checks()


def test_phj_plans_and_hash_seeds_do_not_change_the_result(eng):
    rk, rv, sk, sv, _, _ = oracle_generate(1 << 17, 1 << 18, threads=2, seed=11)
    want = oracle_join("phj", rk, rv, sk, sv, threads=2)
    for opts in ({"radix_bits": (4,)}, {"radix_bits": (8, 8)}, {"radix_bits": (3, 4, 2)}, {"radix_bits": (11, 5)}, {"radix_bits": (9, 7)}, {"radix_bits": (9, 9)},
                 {"radix_bits": (2, 2, 2, 2)}, {"part_tuples": 100}, {"part_tuples": 1 << 20}, {"seed": 12345},
                 {"npj_load": 0.9}):
        assert_same(run(eng, "phj", rk, rv, sk, sv, **opts), want)
        assert_same(run(eng, "npj", rk, rv, sk, sv, **opts), want, rows=False)


def test_misaligned_device_columns_are_rejected(eng):
    k = dev(np.arange(1, 200, dtype=np.uint32))
    with pytest.raises(hj.HjbError):
        eng.npj((k[1:101], k[1:101]), (k[:100], k[:100]))


# ---------------------------------------------------------------- single kernels vs oracle

@pytest.mark.parametrize("bits", [1, 6, 8, 11])
def test_histogram_kernel_matches_oracle(eng, bits):
    rk, rv, sk, sv, _, _ = oracle_generate(1000, 300001, threads=2, seed=13)
    f = eng.hash_factor(0, 0)
    want = np.zeros(1 << bits, np.uint32)
    olib().hjo_histogram(_p(sk), sk.size, _p(want), f, 1 << bits)     # h(key, f, 2^bits), cpra2.cpp:730-741
    assert (eng.histogram(dev(sk), f, 0, bits) == want).all()


def _canon(k, v, off):                   # order inside a partition is not part of the contract
    k, v = k.copy(), v.copy()
    for p in range(off.size - 1):
        s = slice(int(off[p]), int(off[p + 1]))
        o = np.lexsort((v[s], k[s]))
        k[s], v[s] = k[s][o], v[s][o]
    return k, v


def test_partition_kernels_reproduce_the_reference_s_own_partitions(eng):
    """faithful factors: under the hash factor the reference ran with, the CUDA histogram and
    scatter must give the counts and the partition contents the reference's compiled histogram() /
    partition() produced (tests/golden/ref_vectors.npz) -- 64 partitions in one pass, 4096 in two"""
    sk, sv, f = G["in_small_sk"], G["in_small_sv"], int(G["part_factor"])
    assert (eng.histogram(dev(sk), f, 0, 6) == G["ref_small_hist_64"]).all()
    k1, v1, off1 = eng.partition_pass(dev(sk), dev(sv), f, 0, 6)
    assert (off1 == np.concatenate([[0], np.cumsum(G["ref_small_hist_64"])])).all()
    got = _canon(k1.cpu().numpy().view(np.uint32), v1.cpu().numpy().view(np.uint32), off1)
    ref = _canon(G["ref_small_part_keys_64"], G["ref_small_part_vals_64"], off1)
    assert (got[0] == ref[0]).all() and (got[1] == ref[1]).all()
    k2, v2, off2 = eng.partition_pass(k1, v1, f, 6, 6, parent_offsets=off1)
    assert (off2 == np.concatenate([[0], np.cumsum(G["ref_small_hist_4096"])])).all()
    got = _canon(k2.cpu().numpy().view(np.uint32), v2.cpu().numpy().view(np.uint32), off2)
    ref = _canon(G["ref_small_part_keys_4096"], G["ref_small_part_vals_4096"], off2)
    assert (got[0] == ref[0]).all() and (got[1] == ref[1]).all()


def test_histogram_under_mt19937_drawn_factors(eng):
    """the reference draws a fresh odd factor per pass from its MT19937 stream (cpra2.cpp:1787)"""
    import ctypes as C

    class R32(C.Structure):
        _fields_ = [("num", C.c_uint32 * 625), ("index", C.c_size_t)]
    L = olib()
    L.hjo_rand32_next.restype = C.c_uint32
    st = R32()
    L.hjo_rand32_seed(C.byref(st), C.c_uint32(2024))
    rk, rv, sk, sv, _, _ = oracle_generate(1000, 200003, threads=2, seed=15)
    for bits in (3, 8, 11):
        f = int(L.hjo_rand32_next(C.byref(st))) | 1
        want = np.zeros(1 << bits, np.uint32)
        L.hjo_histogram(_p(sk), sk.size, _p(want), f, 1 << bits)
        assert (eng.histogram(dev(sk), f, 0, bits) == want).all()


@pytest.mark.parametrize("bits1,bits2", [(6, 6), (8, 8), (3, 11), (11, 0), (1, 1)])
def test_partition_pass_kernels_match_oracle(eng, bits1, bits2):
    rk, rv, sk, sv, _, _ = oracle_generate(1000, 500003, threads=2, seed=14)
    f = eng.hash_factor(0, 0)
    P1 = 1 << bits1
    c1 = np.zeros(P1, np.uint32)
    olib().hjo_histogram(_p(sk), sk.size, _p(c1), f, P1)
    wk, wv = np.empty_like(sk), np.empty_like(sv)
    olib().hjo_partition(_p(sk), _p(sv), sk.size, _p(c1), _p(wk), _p(wv), f, P1)
    k1, v1, off1 = eng.partition_pass(dev(sk), dev(sv), f, 0, bits1)
    assert (off1 == np.concatenate([[0], np.cumsum(c1)])).all()

    def canon(k, v, off):                # order inside a partition is not part of the contract
        k, v = k.copy(), v.copy()
        for p in range(off.size - 1):
            s = slice(int(off[p]), int(off[p + 1]))
            o = np.lexsort((v[s], k[s]))
            k[s], v[s] = k[s][o], v[s][o]
        return k, v
    gk, gv = canon(k1.cpu().numpy().view(np.uint32), v1.cpu().numpy().view(np.uint32), off1)
    ok, ov = canon(wk, wv, off1)
    assert (gk == ok).all() and (gv == ov).all()
    if bits2:
        P = 1 << (bits1 + bits2)
        c2 = np.zeros(P, np.uint32)
        olib().hjo_histogram(_p(sk), sk.size, _p(c2), f, P)           # two passes == one pass of bits1+bits2
        k2, v2, off2 = eng.partition_pass(k1, v1, f, bits1, bits2, parent_offsets=off1)
        assert (off2 == np.concatenate([[0], np.cumsum(c2)])).all()
        wk2, wv2 = np.empty_like(sk), np.empty_like(sv)
        olib().hjo_partition(_p(sk), _p(sv), sk.size, _p(c2), _p(wk2), _p(wv2), f, P)
        gk, gv = canon(k2.cpu().numpy().view(np.uint32), v2.cpu().numpy().view(np.uint32), off2)
        ok, ov = canon(wk2, wv2, off2)
        assert (gk == ok).all() and (gv == ov).all()


def test_npj_build_kernel_holds_exactly_the_build_side(eng):
    rk, rv, sk, sv, _, _ = oracle_generate(100000, 100000, threads=2, seed=15)
    rk[:500] = rk[500:1000]
    buckets = 60000
    f = eng.hash_factor(0, 1)
    tab = eng.npj_build(dev(rk), dev(rv), buckets, f).cpu().numpy().view(np.uint64)
    used = tab[tab != np.uint64(0xFFFFFFFFFFFFFFFF)]
    pairs = (rv.astype(np.uint64) << np.uint64(32)) | rk.astype(np.uint64)
    assert (np.sort(used) == np.sort(pairs)).all()
    # chain invariant the probe relies on: between a key's home bucket and its slot every bucket is full
    slot_bucket = np.nonzero(tab != np.uint64(0xFFFFFFFFFFFFFFFF))[0] // 4
    keys = (tab[tab != np.uint64(0xFFFFFFFFFFFFFFFF)] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    home = ((keys * np.uint32(f)).astype(np.uint64) * np.uint64(buckets)) >> np.uint64(32)
    full = (tab.reshape(-1, 4) != np.uint64(0xFFFFFFFFFFFFFFFF)).all(axis=1)
    for h, b in zip(home[home != slot_bucket][:2000], slot_bucket[home != slot_bucket][:2000]):
        h, b = int(h), int(b)
        rng = range(h, b) if b >= h else list(range(h, buckets)) + list(range(0, b))
        assert all(full[x] for x in rng)


# ---------------------------------------------------------------- generator

def test_device_generator_equals_numpy_mirror(eng):
    for kind, n, domain in ((0, 100003, 100003), (1, 300000, 5000), (0, 1 << 16, 1 << 16)):
        k, v = eng.generate(kind, n, domain, 42, 3, datagen.OUTER_FACTOR)
        hk, hv = datagen.generate(kind, n, domain, 42, 3, datagen.OUTER_FACTOR)
        assert (k.cpu().numpy().view(np.uint32) == hk).all() and (v.cpu().numpy().view(np.uint32) == hv).all()
        assert eng.column_sum(k) == int(hk.astype(np.uint64).sum())
    k, _ = eng.generate(0, 1000, 5000, 42, 3, datagen.OUTER_FACTOR, first=2000, total=5000)
    hk, _ = datagen.generate(0, 5000, 5000, 42, 3, datagen.OUTER_FACTOR)
    assert (k.cpu().numpy().view(np.uint32) == hk[2000:3000]).all()


def test_skewed_generator_and_join(eng):
    nr, ns = 1 << 16, 1 << 21
    rk, rv = eng.generate(0, nr, nr, 21, 1, datagen.INNER_FACTOR)
    sk, sv = eng.generate(2, ns, nr, 21, 2, datagen.OUTER_FACTOR, theta=1.0, selectivity=0.5)
    hrk, hsk = rk.cpu().numpy().view(np.uint32), sk.cpu().numpy().view(np.uint32)
    hits = np.isin(hsk, hrk)
    assert abs(hits.mean() - 0.5) < 0.01                          # 50 % of TUPLES match
    top = np.unique(hsk[hits], return_counts=True)[1].max() / hits.sum()
    assert 0.03 < top < 0.12                                      # rank-1 key of a theta=1 law over 2^16 keys
    want = numpy_join(hrk, rv.cpu().numpy().view(np.uint32), hsk, sv.cpu().numpy().view(np.uint32), materialize=False)
    assert want.count == int(hits.sum())
    for algo in ALGOS:
        assert getattr(eng, algo)((rk, rv), (sk, sv)).checks() == want.checks()


# ---------------------------------------------------------------- CPRA on one GPU (G virtual owners)

@pytest.mark.parametrize("G_", [1, 2, 4, 8])
def test_cpra_split_exchange_join_single_process(eng, G_):
    """the multi-GPU data path with the G ranks played one after the other on one device:
    split every chunk by owner, hand each owner its pieces (what the all-to-all does), join
    locally below the owner bits, add up -- must equal the oracle's CPRA on the whole input"""
    nr, ns = 200000, 600000
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=2, seed=16)
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    pieces = {g: {"rk": [], "rv": [], "sk": [], "sv": []} for g in range(G_)}
    for c in range(G_):
        cr, cs = slice(c * nr // G_, (c + 1) * nr // G_), slice(c * ns // G_, (c + 1) * ns // G_)
        sp = eng.cpra_split((dev(rk[cr]), dev(rv[cr])), (dev(sk[cs]), dev(sv[cs])), G_)
        assert sp["r_offsets"][-1] == cr.stop - cr.start and sp["s_offsets"][-1] == cs.stop - cs.start
        for g in range(G_):
            for name, col, off in (("rk", "r_keys", "r_offsets"), ("rv", "r_vals", "r_offsets"),
                                   ("sk", "s_keys", "s_offsets"), ("sv", "s_vals", "s_offsets")):
                pieces[g][name].append(sp[col][sp[off][g]:sp[off][g + 1]].clone())
    total = [0, 0, 0, 0]
    rows = []
    for g in range(G_):
        cols = {n: torch.cat(v) for n, v in pieces[g].items()}
        res = eng.cpra_join_local((cols["rk"], cols["rv"]), (cols["sk"], cols["sv"]), g, G_)
        for i, x in enumerate(res.checks()):
            total[i] = (total[i] + x) & ((1 << 64) - 1)
        rows.append(res.rows_numpy())
    assert tuple(total) == want.checks()
    got_rows = sort_rows(*(np.concatenate([r[i] for r in rows]) for i in range(3)))
    assert (got_rows == want.sorted_rows()).all()


# ---------------------------------------------------------------- BASELINE sizes: size-independent properties

@pytest.mark.parametrize("algo,name", [("phj", "phj_cfg2"), ("npj", "npj_cfg1"), ("npj", "phj_cfg2"), ("phj", "npj_cfg1"),
                                       ("npj", "small_cfg3"), ("phj", "small_cfg3")])
def test_full_size_configs_by_properties(eng, algo, name):
    """BASELINE.json configs 1, 2 and 3 at full size, inputs generated on the device.  Every probe
    key has exactly one build partner, so count = |S| and the checksums are plain column sums of
    S: sum_key = sum(S.key), sum_outer = sum(S.val), sum_inner = sum(S.key * f_R mod 2^32)."""
    nr, ns, kind = datagen.workload(name)
    rk, rv = eng.generate(0, nr, nr, 42, 1, datagen.INNER_FACTOR)
    sk, sv = eng.generate(kind, ns, nr, 42, 2, datagen.OUTER_FACTOR)
    res = getattr(eng, algo)((rk, rv), (sk, sv))
    assert res.count == ns
    assert res.sum_key == eng.column_sum(sk) and res.sum_outer == eng.column_sum(sv)
    inner = (sk.to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF
    assert res.sum_inner == int(inner.sum().item()) & ((1 << 64) - 1)
    k, o, i = res.rows_torch()
    # the rows are a permutation of S extended by the build payload: same multiset of keys ...
    assert eng.column_sum(k) == res.sum_key and eng.column_sum(o) == res.sum_outer and eng.column_sum(i) == res.sum_inner
    # ... every row is internally consistent (payloads are functions of the key)
    kk = k.to(torch.int64) & 0xFFFFFFFF
    assert bool(((kk * datagen.OUTER_FACTOR & 0xFFFFFFFF) == (o.to(torch.int64) & 0xFFFFFFFF)).all())
    assert bool(((kk * datagen.INNER_FACTOR & 0xFFFFFFFF) == (i.to(torch.int64) & 0xFFFFFFFF)).all())
    # ... and idempotence: a second run gives the same checks
    assert getattr(eng, algo)((rk, rv), (sk, sv), materialize=False).checks() == res.checks()


def test_config5_skewed_probe_scaled_against_numpy_and_full_size_by_properties(eng):
    """BASELINE.json config 5 (|R| = 2^27, |S| = 2^30, Zipf theta = 1 probe keys, 50 % of the probe tuples match).
    Scaled by 1/64 the rows are compared with an independent numpy join; at full size NPJ and PHJ -- two unrelated
    algorithms -- must agree on count, checksums and the multiset of rows (device fingerprint), every row must be
    internally consistent and about half of S must match."""
    for shift, exact in ((6, True), (0, False)):
        nr, ns = (1 << 27) >> shift, (1 << 30) >> shift
        rk, rv = eng.generate(0, nr, nr, 42, 1, datagen.INNER_FACTOR)
        sk, sv = eng.generate(2, ns, nr, 42, 2, datagen.OUTER_FACTOR, theta=1.0, selectivity=0.5)
        a = eng.phj((rk, rv), (sk, sv))
        fa, ca = eng.rows_fingerprint(*a.rows_torch()), a.checks()
        k, o, i = a.rows_torch()
        kk = k.to(torch.int64) & 0xFFFFFFFF
        assert bool(((kk * datagen.OUTER_FACTOR & 0xFFFFFFFF) == (o.to(torch.int64) & 0xFFFFFFFF)).all())
        assert bool(((kk * datagen.INNER_FACTOR & 0xFFFFFFFF) == (i.to(torch.int64) & 0xFFFFFFFF)).all())
        del kk, k, o, i
        b = eng.npj((rk, rv), (sk, sv))
        assert b.checks() == ca and eng.rows_fingerprint(*b.rows_torch()) == fa
        assert abs(ca[0] / ns - 0.5) < 0.01
        if exact:
            want = numpy_join(*(t.cpu().numpy().view(np.uint32) for t in (rk, rv, sk, sv)))
            assert ca == want.checks()
            assert (sort_rows(*a.rows_numpy()) == want.sorted_rows()).all()


def test_rows_fingerprint_kernel_equals_numpy_mirror_and_ignores_row_order(eng):
    """the verifier: (sum, xor) of a 64-bit mix of each row -- equal for equal multisets of rows"""
    from hash_join_codes_knl_b200.api import rows_fingerprint_numpy
    rk, rv, sk, sv, _, _ = oracle_generate(30000, 100000, threads=2, seed=41)
    want = oracle_join("npj", rk, rv, sk, sv, threads=2)
    fp_oracle = rows_fingerprint_numpy(*want.rows)
    for algo in ALGOS:
        got = run(eng, algo, rk, rv, sk, sv)
        assert eng.rows_fingerprint(*got.rows_torch()) == fp_oracle
    k, o, i = (dev(c) for c in want.rows)
    perm = torch.randperm(k.numel(), device=k.device)
    assert eng.rows_fingerprint(k[perm].contiguous(), o[perm].contiguous(), i[perm].contiguous()) == fp_oracle
    o2 = o.clone()
    o2[17] ^= 1                                                 # one flipped bit in one row
    assert eng.rows_fingerprint(k, o2, i) != fp_oracle
    assert eng.rows_fingerprint(k[:0], o[:0], i[:0]) == (0, 0)


def test_full_size_config2_npj_and_phj_emit_the_same_multiset_of_rows(eng):
    """BASELINE config 2 at full size (2^27 x 2^27): 2^27 rows each from two different algorithms,
    compared as multisets through the device fingerprint -- no copy, no sort"""
    n = 1 << 27
    rk, rv = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
    sk, sv = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
    a = eng.phj((rk, rv), (sk, sv))
    fa, ca = eng.rows_fingerprint(*a.rows_torch()), a.checks()
    b = eng.npj((rk, rv), (sk, sv))
    fb, cb = eng.rows_fingerprint(*b.rows_torch()), b.checks()
    assert ca == cb and ca[0] == n and fa == fb
    # the same rows rebuilt from S alone (every probe tuple has exactly one partner, payloads are functions of the key)
    inner = ((sk.to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF).to(torch.int32)
    assert eng.rows_fingerprint(sk, sv, inner) == fa


def test_cpra_join_local_rejects_foreign_tuples(eng):
    """the local join trusts that every tuple hashes into its owner's range (the DIRECT tables rely
    on it); a caller that hands it somebody else's tuples gets an error, not a wrong result"""
    nr, ns = 1 << 21, 1 << 21
    rk, rv = eng.generate(0, nr, nr, 5, 1, datagen.INNER_FACTOR)
    sk, sv = eng.generate(0, ns, nr, 5, 2, datagen.OUTER_FACTOR)
    with pytest.raises(hj.HjbError):
        eng.cpra_join_local((rk, rv), (sk, sv), 1, 4)           # all owners' tuples, claimed to be owner 1's


# ---------------------------------------------------------------- the reference's command line

def test_cli_programs_keep_the_reference_interface(tmp_path):
    """./write then ./npj | ./phj | ./cpra [#threads] [outer] [inner] on the four raw uint32 files
    (write.cpp:1824-1865, npj.cpp:929-1039): first stdout line is the reference's seconds line, the
    JSON line after it must carry the oracle's count and checksums for the files written."""
    import json
    import subprocess
    from hash_join_codes_knl_b200 import api, build
    build.build_programs()
    bin_dir = os.path.join(os.path.dirname(os.path.abspath(hj.__file__)), "bin")
    nr, ns = 50000, 200000
    subprocess.check_call([os.path.join(bin_dir, "write"), "4", str(ns), str(nr)], cwd=tmp_path)
    assert sorted(os.listdir(tmp_path)) == [f"ik_{nr}.txt", f"iv_{nr}.txt", f"ok_{ns}.txt", f"ov_{ns}.txt"]
    rk, rv = api.relation_read(tmp_path, False, nr)
    sk, sv = api.relation_read(tmp_path, True, ns)
    assert np.unique(rk).size == nr and (rk != 0).all() and np.isin(sk, rk).all()
    want = oracle_join("npj", rk, rv, sk, sv, threads=2, materialize=False)
    for prog in ("npj", "phj", "cpra"):
        out = subprocess.run([os.path.join(bin_dir, prog), "4", str(ns), str(nr), "1"], cwd=tmp_path,
                             capture_output=True, text=True, check=True).stdout.strip().splitlines()
        first = out[1] if prog == "cpra" else out[0]
        assert out[0].startswith("copy:") if prog == "cpra" else True
        float(first.split()[0])                                   # the reference's "%.4f" / "%lf" seconds line
        js = json.loads(out[-1])
        assert (js["join_tuples"], js["sum_key"], js["sum_outer"], js["sum_inner"]) == want.checks(), prog
    # selectivity / skew knobs of ./write (write.cpp:1685-1686), which the reference accepts but does not honour
    subprocess.check_call([os.path.join(bin_dir, "write"), "4", str(ns), str(nr), "0.5", "1.0"], cwd=tmp_path)
    sk2, _ = api.relation_read(tmp_path, True, ns)
    assert abs(np.isin(sk2, rk).mean() - 0.5) < 0.02
    with pytest.raises(subprocess.CalledProcessError):
        subprocess.run([os.path.join(bin_dir, "npj"), "4", "12345", str(nr)], cwd=tmp_path, check=True,
                       capture_output=True)                        # no such file: an error, not garbage
