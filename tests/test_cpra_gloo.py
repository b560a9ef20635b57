"""world_size-2 (and 4) gloo run of the CPRA exchange plumbing on CPU tensors: owner split ->
count exchange -> variable all-to-all -> local join -> all-reduce of the checksums.  The GPU
pieces (split, local join) are stood in for by numpy + the oracle HERE ONLY, so that the
torch.distributed logic of hash_join_codes_knl_b200/cpra.py is covered without a GPU; the
product path always runs them on the Engine (tests/test_join_gpu.py, bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

FACTOR = 0x9E3779B1


def _worker(rank, world, port, nr, ns, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _oracle import oracle_generate, oracle_join
    from hash_join_codes_knl_b200 import cpra
    rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=2, seed=5)
    want = oracle_join("cpra", rk, rv, sk, sv, threads=2, materialize=False)
    # this rank's contiguous chunk (cpra2.cpp:1724-1731)
    cr = slice(rank * nr // world, (rank + 1) * nr // world)
    cs = slice(rank * ns // world, (rank + 1) * ns // world)
    gbits = world.bit_length() - 1

    def split_fn(inner, outer, g):
        out = {}
        for name, (k, v) in (("r", inner), ("s", outer)):
            k, v = k.numpy().view(np.uint32), v.numpy().view(np.uint32)
            owner = ((k * np.uint32(FACTOR)) >> np.uint32(32 - gbits)).astype(np.int64) if gbits else np.zeros(k.size, np.int64)
            order = np.argsort(owner, kind="stable")
            out[name + "_keys"] = torch.from_numpy(k[order].view(np.int32).copy())
            out[name + "_vals"] = torch.from_numpy(v[order].view(np.int32).copy())
            out[name + "_offsets"] = [0] + list(np.cumsum(np.bincount(owner, minlength=g)))
        return out

    def join_fn(inner, outer, me, g):
        k = inner[0].numpy().view(np.uint32)
        owner = (k * np.uint32(FACTOR)) >> np.uint32(32 - gbits) if gbits else np.zeros(k.size, np.uint32)
        assert (owner == me).all()                       # every received tuple belongs to this owner
        return oracle_join("npj", k, inner[1].numpy().view(np.uint32), outer[0].numpy().view(np.uint32),
                           outer[1].numpy().view(np.uint32), materialize=False)

    t = lambda a: torch.from_numpy(a.view(np.int32).copy())
    res = cpra.cpra_join(None, (t(rk[cr]), t(rv[cr])), (t(sk[cs]), t(sv[cs])), split_fn=split_fn, join_fn=join_fn)
    got = (res["count"], res["sum_key"], res["sum_outer"], res["sum_inner"])
    # conservation: what arrived everywhere is what was sent
    n_recv = torch.tensor(list(res["recv_tuples"]), dtype=torch.int64)
    dist.all_reduce(n_recv)
    ok = got == want.checks() and n_recv.tolist() == [nr, ns]
    ret[rank] = (ok, got, want.checks())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cpra_exchange_gloo(world):
    from _oracle import build_oracle
    build_oracle()
    port = 29500 + (os.getpid() % 2000) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, 20000, 60000, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, got, want = ret[r]
        assert ok, (r, got, want)


def test_reduce_checks_wraps_like_uint64():
    from hash_join_codes_knl_b200.cpra import reduce_checks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(31000 + os.getpid() % 2000)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        big = (1 << 64) - 5
        assert reduce_checks(big, 1, 2, (1 << 63) + 7, torch.device("cpu")) == (big, 1, 2, (1 << 63) + 7)
    finally:
        dist.destroy_process_group()
