"""The product exchange path of CPRA -- hjb_cpra_count -> hjb_cpra_scatter_peer (k_scatter_bulk: TMA bulk
copies into the owners' receive buffers) -> hjb_cpra_join_local, and its stream-ordered form hjb_cpra_bind /
_count_async / _scatter_async / _join_async / _finish -- against the oracle's CPRA (oracle/hj_oracle.c,
cpra2.cpp:1697-1986).  On ONE GPU the G owners are G contexts whose receive buffers are local allocations
(the kernel cannot tell a local column from a peer-mapped one); with >= 2 GPUs the same path runs under
torchrun with CUDA IPC and NCCL (tests/test_cpra_nccl.py).  Reference step replaced: the per-owner gather
cpra2.cpp:1861-1905,1940-1959."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
from hash_join_codes_knl_b200.api import HjbCapacityError
from _oracle import numpy_join, oracle_generate, oracle_join, sort_rows

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MASK64 = (1 << 64) - 1


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


@pytest.fixture(scope="module")
def engines():
    es = [hj.Engine(0) for _ in range(8)]
    yield es
    for e in es:
        e.close()


def chunk(a, c, G):
    return a[c * a.size // G:(c + 1) * a.size // G]


def recv_buffers(G, r_cap, s_cap):
    """per owner: r_keys r_vals s_keys s_vals, filled with a pattern no tuple carries, so rows that were never
    written show up as wrong results"""
    bufs = [[torch.full((cap + 64,), -559038737, dtype=torch.int32, device="cuda") for cap in (r_cap, r_cap, s_cap, s_cap)]
            for _ in range(G)]
    peers = [[bufs[g][c].data_ptr() for g in range(G)] for c in range(4)]
    torch.cuda.synchronize()                 # the contexts run on their own non-blocking streams
    return bufs, peers


def add_checks(total, res):
    return [(a + b) & MASK64 for a, b in zip(total, res.checks())]


def run_sync_path(engines, G, rk, rv, sk, sv):
    """count on every sender -> bases from the count matrix (host) -> scatter into the owners' buffers -> local joins"""
    counts = []
    torch.cuda.synchronize()
    for c in range(G):
        counts.append(engines[c].cpra_count((dev(chunk(rk, c, G)), dev(chunk(rv, c, G))),
                                            (dev(chunk(sk, c, G)), dev(chunk(sv, c, G))), G))
    r_recv = [sum(counts[s][0][g] for s in range(G)) for g in range(G)]
    s_recv = [sum(counts[s][1][g] for s in range(G)) for g in range(G)]
    assert sum(r_recv) == rk.size and sum(s_recv) == sk.size
    bufs, peers = recv_buffers(G, max(r_recv), max(s_recv))
    for c in range(G):
        r_base = [sum(counts[s][0][g] for s in range(c)) for g in range(G)]
        s_base = [sum(counts[s][1][g] for s in range(c)) for g in range(G)]
        engines[c].cpra_scatter_peer(G, peers, r_base, s_base)
    total, rows = [0, 0, 0, 0], []
    for g in range(G):
        res = engines[g].cpra_join_local((bufs[g][0][:r_recv[g]], bufs[g][1][:r_recv[g]]),
                                         (bufs[g][2][:s_recv[g]], bufs[g][3][:s_recv[g]]), g, G)
        total = add_checks(total, res)
        rows.append(res.rows_numpy())
    return tuple(total), rows, (r_recv, s_recv)


def run_async_path(engines, G, rk, rv, sk, sv, r_cap, s_cap, hot_keys=None):
    """the stream-ordered step; the all-gather is a torch.cat, the cross-GPU ordering a device synchronise.
    hot_keys: the probe tuples with these keys stay on their sender and are joined there (skew handling)"""
    bufs, peers = recv_buffers(G, r_cap, s_cap)
    counts = [torch.zeros(2 * G, dtype=torch.int64, device="cuda") for _ in range(G)]
    keep, hot_s, hot_r = [], [], None
    for c in range(G):
        engines[c].cpra_bind(c, G, peers, r_cap, s_cap)
        cols = ((dev(chunk(rk, c, G)), dev(chunk(rv, c, G))), (dev(chunk(sk, c, G)), dev(chunk(sv, c, G))))
        keep.append(cols)
    if hot_keys is not None:
        hk = dev(np.sort(np.asarray(hot_keys, np.uint32)))
        parts_k, parts_v = [], []
        for c in range(G):
            ok_, ov_ = torch.zeros(4096, dtype=torch.int32, device="cuda"), torch.zeros(4096, dtype=torch.int32, device="cuda")
            found = engines[c].cpra_select_hot(keep[c][0], hk, ok_, ov_)
            parts_k.append(ok_[:found].clone())
            parts_v.append(ov_[:found].clone())
            cold, hot = engines[c].cpra_split_hot(keep[c][1], hk)
            assert cold[0].numel() + hot[0].numel() == keep[c][1][0].numel()
            keep[c] = (keep[c][0], cold)
            hot_s.append(hot)
        hot_r = (torch.cat(parts_k).contiguous(), torch.cat(parts_v).contiguous())
    torch.cuda.synchronize()
    for c in range(G):
        engines[c].cpra_count_async(keep[c][0], keep[c][1], counts[c])
    torch.cuda.synchronize()
    matrix = torch.cat(counts).contiguous()
    torch.cuda.synchronize()
    for c in range(G):
        engines[c].cpra_scatter_async(matrix)
    torch.cuda.synchronize()
    total, rows, recv = [0, 0, 0, 0], [], []
    err = None
    for g in range(G):
        engines[g].cpra_join_async()
        if hot_r is not None:
            engines[g].cpra_hot_join(hot_s[g], hot_r)
        try:
            res, got, largest = engines[g].cpra_finish()
        except HjbCapacityError as e:
            err = e
            continue
        total = add_checks(total, res)
        rows.append(res.rows_numpy())
        recv.append(got)
    if err is not None:
        raise err
    return tuple(total), rows, recv, matrix.view(G, 2 * G).cpu().numpy()


def all_rows(rows):
    return sort_rows(*(np.concatenate([r[i] for r in rows]) for i in range(3)))


def skewed(nr, ns, seed):
    """a third of the probe side is one heavy-hitter key (all of it lands on one owner), a tenth has no partner"""
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=2, seed=seed)
    sk = sk.copy()
    sk[::3] = rk[7]
    sk[5::10] ^= np.uint32(0x10000000)
    return rk, rv, sk, sv


CASES = [("uniform", 200000, 600000, 16), ("skewed", 150000, 500000, 19), ("tiny", 37, 5, 20), ("few", 4096, 33, 21),
         ("one", 1, 70000, 22), ("ragged", 100003, 70001, 18)]


@pytest.mark.parametrize("G_", [2, 4, 8])
@pytest.mark.parametrize("name,nr,ns,seed", CASES)
def test_fused_exchange_with_virtual_owners_matches_oracle(engines, G_, name, nr, ns, seed):
    rk, rv, sk, sv = skewed(nr, ns, seed) if name == "skewed" else oracle_generate(nr, ns, threads=2, seed=seed)[:4]
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    got, rows, _ = run_sync_path(engines, G_, rk, rv, sk, sv)
    assert got == want.checks()
    assert (all_rows(rows) == want.sorted_rows()).all()


@pytest.mark.parametrize("G_", [2, 4, 8])
@pytest.mark.parametrize("name,nr,ns,seed", CASES)
def test_stream_ordered_step_with_virtual_owners_matches_oracle(engines, G_, name, nr, ns, seed):
    rk, rv, sk, sv = skewed(nr, ns, seed) if name == "skewed" else oracle_generate(nr, ns, threads=2, seed=seed)[:4]
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    got, rows, recv, matrix = run_async_path(engines, G_, rk, rv, sk, sv, nr + 1024, ns + 1024)
    assert got == want.checks()
    assert (all_rows(rows) == want.sorted_rows()).all()
    assert [r[0] for r in recv] == [int(matrix[:, g].sum()) for g in range(G_)]
    assert [r[1] for r in recv] == [int(matrix[:, G_ + g].sum()) for g in range(G_)]


def test_stream_ordered_step_reports_a_receive_buffer_that_is_too_small(engines):
    """the owner of the heavy hitter receives far more than a uniform share: every sender sees it in the count
    matrix, nothing is scattered, finish raises with the size the fullest owner needs; a second step with
    that capacity gives the oracle's rows"""
    rk, rv, sk, sv = skewed(150000, 500000, 23)
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    G_ = 4
    with pytest.raises(HjbCapacityError) as info:
        run_async_path(engines, G_, rk, rv, sk, sv, rk.size // G_ + 4096, sk.size // G_ + 4096)
    need_r, need_s = info.value.largest
    assert need_s > sk.size // 3 and need_r >= rk.size // G_ - 4096
    got, rows, _, _ = run_async_path(engines, G_, rk, rv, sk, sv, need_r, need_s)
    assert got == want.checks() and (all_rows(rows) == want.sorted_rows()).all()


def test_hot_keys_stay_with_their_sender_and_the_owners_receive_balanced_shares(engines):
    """skew handling (hjb_cpra_split_hot / _select_hot / _hot_join): a third of the probe side is ONE key, another
    key takes a tenth and has no partner; with the frequent keys declared hot the rows are still the oracle's, and no owner
    receives much more than its share"""
    G_ = 4
    rk, rv, sk, sv = skewed(150000, 500000, 24)
    sk = sk.copy()
    sk[1::10] = np.uint32(0x12345679)                       # frequent, but not a build key
    assert not (rk == np.uint32(0x12345679)).any()
    rk[11], rv[11] = rk[7], np.uint32(99)                    # the hot key twice on the build side: every pair is emitted
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    _, _, recv_plain, _ = run_async_path(engines, G_, rk, rv, sk, sv, rk.size, sk.size)
    hot = [rk[7], 0x12345679, rk[7] ^ np.uint32(0x10000000)]    # skewed() also turns every 30th tuple into rk[7] ^ 0x10000000
    got, rows, recv, _ = run_async_path(engines, G_, rk, rv, sk, sv, rk.size, sk.size, hot_keys=hot)
    assert got == want.checks() and (all_rows(rows) == want.sorted_rows()).all()
    share = lambda r: max(x[1] for x in r) / (sum(x[1] for x in r) / G_)
    assert share(recv_plain) > 2.0 and share(recv) < 1.1


def test_fused_exchange_many_tiles_per_item(engines):
    """2^25 x 2^25 over 4 virtual owners: items of several 8192-tuple tiles, both staging buffers of the
    double-buffered bulk scatter in use, carries across tiles; checked by count, checksums and the
    device fingerprint of the rows against the rows rebuilt from S"""
    G_, n = 4, 1 << 25
    e0 = engines[0]
    rk, rv = e0.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
    sk, sv = e0.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
    e0.synchronize()
    per = n // G_
    counts = [engines[c].cpra_count((rk[c * per:(c + 1) * per], rv[c * per:(c + 1) * per]),
                                    (sk[c * per:(c + 1) * per], sv[c * per:(c + 1) * per]), G_) for c in range(G_)]
    r_recv = [sum(counts[s][0][g] for s in range(G_)) for g in range(G_)]
    s_recv = [sum(counts[s][1][g] for s in range(G_)) for g in range(G_)]
    bufs, peers = recv_buffers(G_, max(r_recv), max(s_recv))
    for c in range(G_):
        engines[c].cpra_scatter_peer(G_, peers, [sum(counts[s][0][g] for s in range(c)) for g in range(G_)],
                                     [sum(counts[s][1][g] for s in range(c)) for g in range(G_)])
    total, fp = [0, 0, 0, 0], [0, 0]
    for g in range(G_):
        res = engines[g].cpra_join_local((bufs[g][0][:r_recv[g]], bufs[g][1][:r_recv[g]]),
                                         (bufs[g][2][:s_recv[g]], bufs[g][3][:s_recv[g]]), g, G_)
        total = add_checks(total, res)
        f = engines[g].rows_fingerprint(*res.rows_torch())
        fp = [(fp[0] + f[0]) & MASK64, fp[1] ^ f[1]]
    inner = ((sk.to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF)
    assert tuple(total) == (n, e0.column_sum(sk), e0.column_sum(sv), int(inner.sum().item()) & MASK64)
    assert tuple(fp) == e0.rows_fingerprint(sk, sv, inner.to(torch.int32))


def run_staged_path(engines, G, rk, rv, sk, sv, r_cap, s_cap, plan=None, own_alloc=False, parts=1):
    """the staged exchange (hjb_cpra_stage_*): stage A locally by owner and sub-partition, TMA copies of whole runs
    into the owners' buffers, one local pass, join; the all-gather is a torch.cat, the cross-GPU ordering a device
    synchronise.  own_alloc: the receive buffers come from hjb_cpra_recv_alloc (as in the multi-process path) -- stage A
    then writes every GPU's own runs straight to their final rows, the others into the staging region behind"""
    if own_alloc:
        owns = [engines[g].cpra_recv_alloc(r_cap, s_cap) for g in range(G)]
        peers = [[owns[g]["ptrs"][c] for g in range(G)] for c in range(4)]
    else:
        bufs, peers = recv_buffers(G, r_cap, s_cap)
    plan = plan or engines[0].cpra_stage_plan(G, max(1, rk.size // G), max(1, sk.size // G))
    assert plan is not None
    abits, bbits, big = plan
    counts = [torch.zeros(2 << abits, dtype=torch.int64, device="cuda") for _ in range(G)]
    keep = []
    for c in range(G):
        engines[c].cpra_bind(c, G, peers, r_cap, s_cap)
        keep.append(((dev(chunk(rk, c, G)), dev(chunk(rv, c, G))), (dev(chunk(sk, c, G)), dev(chunk(sv, c, G)))))
    torch.cuda.synchronize()
    parts = min(parts, 1 << abits >> (G.bit_length() - 1))
    for c in range(G):
        engines[c].cpra_stage_count_async(keep[c][0], keep[c][1], abits, counts[c], nparts=parts)
    torch.cuda.synchronize()
    matrix = torch.cat(counts).contiguous()
    torch.cuda.synchronize()
    for c in range(G):
        engines[c].cpra_stage_scatter_async(matrix, 0)
        engines[c].cpra_stage_scatter_async(matrix, 1)
    torch.cuda.synchronize()
    # the pieces in the order cpra_join_staged sends them; the owners pass and join piece k while later pieces are still due
    order = [(0, 0)] + [p for k in range(parts) for p in ([(0, k + 1)] if k + 1 < parts else []) + [(1, k)]]
    for i, (rel, k) in enumerate(order):
        for c in range(G):
            engines[c].cpra_stage_copy_async(rel, part=k)
        torch.cuda.synchronize()
        if rel == 1:
            for g in range(G):
                engines[g].cpra_stage_local_async(bbits, big, 0, part=k)
                engines[g].cpra_stage_local_async(bbits, big, 1, part=k)
    torch.cuda.synchronize()
    total, rows, recv = [0, 0, 0, 0], [], []
    err = None
    for g in range(G):
        try:
            res, got, largest = engines[g].cpra_finish()
        except HjbCapacityError as e:
            err = e
            continue
        total = add_checks(total, res)
        rows.append(res.rows_numpy())
        recv.append(got)
    if err is not None:
        raise err
    return tuple(total), rows, recv, matrix.view(G, 2, 1 << abits).cpu().numpy()


@pytest.mark.parametrize("copy", ["tma", "ce"])
@pytest.mark.parametrize("own_alloc", [False, True])
@pytest.mark.parametrize("G_", [2, 4, 8])
@pytest.mark.parametrize("name,nr,ns,seed", CASES)
def test_staged_exchange_with_virtual_owners_matches_oracle(engines, G_, name, nr, ns, seed, own_alloc, copy, monkeypatch):
    """copy: the runs leave through k_peer_copy (TMA bulk copies) or through the copy engines (HJB_STAGE_COPY=ce)"""
    monkeypatch.setenv("HJB_STAGE_COPY", copy)
    rk, rv, sk, sv = skewed(nr, ns, seed) if name == "skewed" else oracle_generate(nr, ns, threads=2, seed=seed)[:4]
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    got, rows, recv, matrix = run_staged_path(engines, G_, rk, rv, sk, sv, nr + 1024, ns + 1024, own_alloc=own_alloc, parts=4 if copy == "ce" else 1)
    assert got == want.checks()
    assert (all_rows(rows) == want.sorted_rows()).all()
    per_owner = matrix.reshape(G_, 2, G_, -1).sum(axis=(0, 3))          # [rel][owner]
    assert [r[0] for r in recv] == [int(x) for x in per_owner[0]]
    assert [r[1] for r in recv] == [int(x) for x in per_owner[1]]


@pytest.mark.parametrize("parts", [1, 2, 8])
@pytest.mark.parametrize("G_,plan", [(2, (9, 9, 1)), (8, (9, 9, 1)), (4, (9, 1, 0)), (2, (2, 1, 0)), (4, (2, 9, 0)), (8, (3, 5, 0)),
                                     (2, (8, 8, 0)), (4, (5, 7, 0))])
def test_staged_exchange_under_explicit_plans(engines, G_, plan, parts):
    """512-way stage A, 12288-tuple join fills, a one-bit local pass, stage A with the owner bits only"""
    rk, rv, sk, sv = skewed(150000, 500000, 31)
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    got, rows, _, _ = run_staged_path(engines, G_, rk, rv, sk, sv, rk.size, sk.size, plan=plan, own_alloc=G_ != 4, parts=parts)
    assert got == want.checks()
    assert (all_rows(rows) == want.sorted_rows()).all()


@pytest.mark.parametrize("own_alloc", [False, True])
def test_staged_exchange_reports_a_receive_buffer_that_is_too_small(engines, own_alloc):
    rk, rv, sk, sv = skewed(150000, 500000, 23)
    want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
    G_ = 4
    with pytest.raises(HjbCapacityError) as info:
        run_staged_path(engines, G_, rk, rv, sk, sv, rk.size // G_ + 4096, sk.size // G_ + 4096, own_alloc=own_alloc)
    need_r, need_s = info.value.largest
    assert need_s > sk.size // 3 and need_r >= rk.size // G_ - 4096
    got, rows, _, _ = run_staged_path(engines, G_, rk, rv, sk, sv, need_r, need_s, own_alloc=own_alloc)
    assert got == want.checks() and (all_rows(rows) == want.sorted_rows()).all()


def test_staged_exchange_regrows_the_result_when_equal_build_keys_multiply_the_rows(engines):
    """the heavy-hitter probe key occurs eight times on the build side: more rows than the result columns were sized for;
    hjb_cpra_finish grows them and runs the join phase again over all parts' partitions"""
    rk, rv, sk, sv = skewed(150000, 500000, 27)
    rk, rv = rk.copy(), rv.copy()
    rk[20:27] = rk[7]
    rv[20:27] = np.arange(7, dtype=np.uint32) + np.uint32(1000)
    want = numpy_join(rk, rv, sk, sv)
    assert want.count > 2 * sk.size
    got, rows, _, _ = run_staged_path(engines, 4, rk, rv, sk, sv, rk.size, sk.size, own_alloc=True, parts=2)
    assert got == want.checks() and (all_rows(rows) == want.sorted_rows()).all()


@pytest.mark.parametrize("plan,parts", [(None, 1), ((9, 9, 1), 4), (None, 4)])
def test_staged_exchange_many_pieces_per_run(engines, plan, parts):
    """2^25 x 2^25 over 4 virtual owners: runs of many 16 KB pieces, every copy stage in use; checked by count,
    checksums and the device fingerprint of the rows against the rows rebuilt from S"""
    G_, n = 4, 1 << 25
    e0 = engines[0]
    rk, rv = e0.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
    sk, sv = e0.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
    e0.synchronize()
    per = n // G_
    plan = plan or e0.cpra_stage_plan(G_, per, per)
    abits, bbits, big = plan
    cap = per + per // 8
    owns = [engines[g].cpra_recv_alloc(cap, cap) for g in range(G_)]
    peers = [[owns[g]["ptrs"][c] for g in range(G_)] for c in range(4)]
    counts = [torch.zeros(2 << abits, dtype=torch.int64, device="cuda") for _ in range(G_)]
    for c in range(G_):
        engines[c].cpra_bind(c, G_, peers, cap, cap)
        engines[c].cpra_stage_count_async((rk[c * per:(c + 1) * per], rv[c * per:(c + 1) * per]),
                                          (sk[c * per:(c + 1) * per], sv[c * per:(c + 1) * per]), abits, counts[c], nparts=parts)
    torch.cuda.synchronize()
    matrix = torch.cat(counts).contiguous()
    torch.cuda.synchronize()
    for rel in (0, 1):
        for c in range(G_):
            engines[c].cpra_stage_scatter_async(matrix, rel)
    torch.cuda.synchronize()
    for k in range(parts):
        for c in range(G_):
            engines[c].cpra_stage_copy_async(0, part=k)
            engines[c].cpra_stage_copy_async(1, part=k)
        torch.cuda.synchronize()
        for g in range(G_):
            engines[g].cpra_stage_local_async(bbits, big, 0, part=k)
            engines[g].cpra_stage_local_async(bbits, big, 1, part=k)
    torch.cuda.synchronize()
    total, fp = [0, 0, 0, 0], [0, 0]
    for g in range(G_):
        res, _, _ = engines[g].cpra_finish()
        total = add_checks(total, res)
        f = engines[g].rows_fingerprint(*res.rows_torch())
        fp = [(fp[0] + f[0]) & MASK64, fp[1] ^ f[1]]
    inner = ((sk.to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF)
    assert tuple(total) == (n, e0.column_sum(sk), e0.column_sum(sv), int(inner.sum().item()) & MASK64)
    assert tuple(fp) == e0.rows_fingerprint(sk, sv, inner.to(torch.int32))


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_cpra_over_nccl_and_cuda_ipc(nproc):
    """one process per GPU, NCCL collectives and peer-mapped receive buffers: tests/test_cpra_nccl.py under torchrun"""
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, HJB_CPRA_RANDOM="6")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29540 + nproc),
                          os.path.join(ROOT, "tests", "test_cpra_nccl.py")], capture_output=True, text=True, timeout=1500, env=env)
    assert out.returncode == 0 and "CPRA_NCCL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
