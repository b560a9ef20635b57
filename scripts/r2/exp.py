"""One line per (config, algorithm): best-of-N device time and per-kernel times, result verified.
usage: exp.py cfg[,cfg...] [algo[,algo]] [materialize 0|1]     cfg in cfg1 cfg2 cfg3 cfg5 cfg4h (2^28 x 2^28)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
cfgs = sys.argv[1].split(",")
algos = sys.argv[2].split(",") if len(sys.argv) > 2 else ["npj", "phj"]
mat = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("HJB_"))
eng = hj.Engine(0)
M = (1 << 64) - 1
SH = {"cfg1": (24, 28, 1), "cfg2": (27, 27, 0), "cfg3": (16, 30, 1), "cfg4h": (28, 28, 0), "cfg5": (27, 30, 2)}
for c in cfgs:
    lr, ls, kind = SH[c]
    R = eng.generate(0, 1 << lr, 1 << lr, 42, 1, datagen.INNER_FACTOR)
    S = eng.generate(kind, 1 << ls, 1 << lr, 42, 2, datagen.OUTER_FACTOR, theta=1.0, selectivity=0.5) if kind == 2 else \
        eng.generate(kind, 1 << ls, 1 << lr, 42, 2, datagen.OUTER_FACTOR)
    want = None
    if kind != 2:
        inner = (S[0].to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF
        want = (S[0].numel(), eng.column_sum(S[0]), eng.column_sum(S[1]), int(inner.sum().item()) & M)
        del inner
    ref = None
    for a in algos:
        eng.set_profiling(False)
        best = None
        for _ in range(4):
            r = getattr(eng, a)(R, S, materialize=bool(mat))
            best = r.seconds if best is None else min(best, r.seconds)
        eng.set_profiling(True)
        getattr(eng, a)(R, S, materialize=bool(mat))
        r = getattr(eng, a)(R, S, materialize=bool(mat))
        kt = {k: round(v[0], 3) for k, v in eng.kernel_times().items() if v[1]}
        ok = "unchecked"
        if want is not None:
            ok = "OK" if r.checks() == want else f"MISMATCH {r.checks()} {want}"
        else:
            ref = ref or r.checks()
            ok = "agree" if r.checks() == ref else "DISAGREE"
        n = (1 << lr) + (1 << ls)
        print(f"[{tag}] {c} {a} mat={mat} {best*1e3:8.3f} ms {n/best/1e9:7.1f} Gt/s {ok} {kt}", flush=True)
    del R, S
    torch.cuda.empty_cache()
