"""the staged exchange on ONE GPU with two virtual owners (2^27 + 2^27 tuples each): per-kernel times of a sender"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
G, n = int(os.environ.get("STAGE_G", "2")), 1 << int(os.environ.get("STAGE_LOG2", "27"))
PLAN = tuple(int(x) for x in os.environ["STAGE_PLAN"].split(",")) if os.environ.get("STAGE_PLAN") else None
es = [hj.Engine(0) for _ in range(G)]
tot = n * G
cols = []
for c in range(G):
    R = es[c].generate(0, n, tot, 42, 1, datagen.INNER_FACTOR, first=c * n, total=tot)
    S = es[c].generate(0, n, tot, 42, 2, datagen.OUTER_FACTOR, first=c * n, total=tot)
    cols.append((R, S))
cap = n + n // 8
for own_alloc in (True, False):
    if own_alloc:
        owns = [es[g].cpra_recv_alloc(cap, cap) for g in range(G)]
        peers = [[owns[g]["ptrs"][c] for g in range(G)] for c in range(4)]
    else:
        bufs = [[torch.empty(cap + 64, dtype=torch.int32, device="cuda") for _ in range(4)] for _ in range(G)]
        peers = [[bufs[g][c].data_ptr() for g in range(G)] for c in range(4)]
    plan = PLAN or es[0].cpra_stage_plan(G, n, n)
    abits, bbits, big = plan
    counts = [torch.zeros(2 << abits, dtype=torch.int64, device="cuda") for _ in range(G)]
    for c in range(G):
        es[c].cpra_bind(c, G, peers, cap, cap)
        es[c].set_profiling(True)
    for it in range(3):
        torch.cuda.synchronize()
        for c in range(G):
            es[c].cpra_stage_count_async(cols[c][0], cols[c][1], abits, counts[c])
        torch.cuda.synchronize()
        matrix = torch.cat(counts).contiguous()
        torch.cuda.synchronize()
        for c in range(G):
            es[c].cpra_stage_scatter_async(matrix, 0)
            es[c].cpra_stage_scatter_async(matrix, 1)
            torch.cuda.synchronize()
        for c in range(G):
            es[c].cpra_stage_copy_async(0)
            es[c].cpra_stage_copy_async(1)
            torch.cuda.synchronize()          # one sender at a time
        total = 0
        for g in range(G):
            es[g].cpra_stage_local_async(bbits, big, 0)
            es[g].cpra_stage_local_async(bbits, big, 1)
            res, got, _ = es[g].cpra_finish()
            total += res.count
            kt = {k: round(v[0], 3) for k, v in es[g].kernel_times().items() if v[1]}
        assert total == tot, (total, tot)
    print("own_alloc", own_alloc, "plan", plan, "owner 1 kernel times (ms):", kt, "copy bytes", (n * 8 if own_alloc else n * 16) / 1e9, "GB", flush=True)
