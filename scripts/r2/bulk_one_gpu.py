"""The peer scatter (k_scatter_bulk) on ONE GPU: 8 virtual owners whose receive buffers are local allocations; an ncu target
(a multi-rank command must not run under ncu).  2^27 + 2^27 tuples, count -> bases -> scatter, twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
G, n = 8, 1 << 27
eng = hj.Engine(0)
R = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
S = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
cap = n // G + n // (4 * G)
bufs = [[torch.empty(cap + 64, dtype=torch.int32, device="cuda") for _ in range(4)] for _ in range(G)]
peers = [[bufs[g][c].data_ptr() for g in range(G)] for c in range(4)]
torch.cuda.synchronize()
for rep in range(2):
    rc, sc = eng.cpra_count(R, S, G)
    ms = eng.cpra_scatter_peer(G, peers, [0] * G, [0] * G)
    print(f"scatter of 2^28 tuples to {G} local owners: {ms:.3f} ms", flush=True)
