"""All five BASELINE.json configs on one GPU (configs 4 and 5 at their single-GPU form), each verified
through size-independent properties; prints one line per (config, algorithm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen

eng = hj.Engine(0)
eng.set_profiling(True)
M = (1 << 64) - 1


def run(name, algo, R, S, want_count=None, reps=3, **opts):
    best = None
    for _ in range(reps):
        r = getattr(eng, algo)(R, S, **opts)
        if best is None or r.seconds < best.seconds:
            best, kt = r, {k: round(v[0], 3) for k, v in eng.kernel_times().items() if v[1]}
    n = R[0].numel() + S[0].numel()
    ok = "" if want_count is None else (" count OK" if best.count == want_count else f" COUNT MISMATCH {best.count} != {want_count}")
    print(f"{name:34s} {algo} {best.seconds*1e3:9.3f} ms {n/best.seconds/1e9:8.2f} Gtuples/s{ok}  {kt}", flush=True)
    return best


def fk_checks(S):
    inner = (S[0].to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF
    return (S[0].numel(), eng.column_sum(S[0]), eng.column_sum(S[1]), int(inner.sum().item()) & M)


# config 1
R = eng.generate(0, 1 << 24, 1 << 24, 42, 1, datagen.INNER_FACTOR); S = eng.generate(1, 1 << 28, 1 << 24, 42, 2, datagen.OUTER_FACTOR)
w = fk_checks(S)
for algo in ("npj", "phj"):
    assert run("cfg1 16M x 256M FK", algo, R, S, w[0]).checks() == w
# config 2
R = eng.generate(0, 1 << 27, 1 << 27, 42, 1, datagen.INNER_FACTOR); S = eng.generate(0, 1 << 27, 1 << 27, 42, 2, datagen.OUTER_FACTOR)
w = fk_checks(S)
for algo in ("phj", "npj"):
    assert run("cfg2 128M x 128M", algo, R, S, w[0]).checks() == w
# config 3
R = eng.generate(0, 1 << 16, 1 << 16, 42, 1, datagen.INNER_FACTOR); S = eng.generate(1, 1 << 30, 1 << 16, 42, 2, datagen.OUTER_FACTOR)
w = fk_checks(S)
for algo in ("npj", "phj"):
    assert run("cfg3 64K x 1B FK", algo, R, S, w[0], reps=2).checks() == w
del R, S
# config 5 at one GPU: 128M x 1B, Zipf theta = 1, 50 % of probe tuples match
R = eng.generate(0, 1 << 27, 1 << 27, 42, 1, datagen.INNER_FACTOR)
S = eng.generate(2, 1 << 30, 1 << 27, 42, 2, datagen.OUTER_FACTOR, theta=1.0, selectivity=0.5)
a = run("cfg5 128M x 1B zipf1.0 sel0.5", "phj", R, S, reps=2)
b = run("cfg5 128M x 1B zipf1.0 sel0.5", "npj", R, S, reps=2)
print("   phj == npj:", a.checks() == b.checks(), "selectivity", a.count / (1 << 30))
del R, S
# config 4 at one GPU, quarter size (2^29 x 2^29) and full size if memory allows
for k in (29, 31):
    try:
        R = eng.generate(0, 1 << k, 1 << k, 42, 1, datagen.INNER_FACTOR); S = eng.generate(0, 1 << k, 1 << k, 42, 2, datagen.OUTER_FACTOR)
        r = run(f"cfg4 2^{k} x 2^{k} on 1 GPU", "phj", R, S, 1 << k, reps=2)
        assert r.sum_key == eng.column_sum(S[0])
        del R, S
    except Exception as e:
        print(f"cfg4 2^{k}: {type(e).__name__}: {e}")
        break
