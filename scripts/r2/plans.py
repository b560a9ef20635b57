"""PHJ config 2 under explicit radix plans: per-kernel times (is a 2048-way pass affordable?)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
eng = hj.Engine(0)
n = 1 << 27
R = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
S = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
for plan in ((8, 8), (5, 11), (8, 11), (11, 5), (11, 8), (7, 8), (8, 7)):
    eng.set_profiling(False)
    for _ in range(2):
        r = eng.phj(R, S, radix_bits=plan)
    eng.set_profiling(True)
    eng.phj(R, S, radix_bits=plan)
    r = eng.phj(R, S, radix_bits=plan)
    kt = {k: round(v[0], 3) for k, v in eng.kernel_times().items() if v[1]}
    print(plan, f"{r.seconds*1e3:.3f} ms count {r.count} phases {[round(x,3) for x in r.phase_ms[:5]]}", kt, flush=True)
