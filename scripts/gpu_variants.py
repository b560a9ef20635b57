"""Times the kernels under option variants on the GPU box (per-kernel CUDA-event times from the
library's profiling facility).  Usage: python scripts/gpu_variants.py [phj|npj|all]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen

which = sys.argv[1] if len(sys.argv) > 1 else "all"
eng = hj.Engine(0)
eng.set_profiling(True)


def run(algo, R, S, reps=3, **opts):
    best = None
    for _ in range(reps):
        r = getattr(eng, algo)(R, S, **opts)
        kt = {k: round(v[0], 3) for k, v in eng.kernel_times().items() if v[1]}
        if best is None or r.seconds < best[0]:
            best = (r.seconds, kt, r.count)
    print(f"{algo:4s} {str(opts):60s} {best[0]*1e3:8.3f} ms  {best[1]}", flush=True)


if which in ("phj", "all"):
    nr = ns = 1 << 27
    R = eng.generate(0, nr, nr, 42, 1, datagen.INNER_FACTOR)
    S = eng.generate(0, ns, nr, 42, 2, datagen.OUTER_FACTOR)
    run("phj", R, S)
    run("phj", R, S, materialize=False)
    run("phj", R, S, radix_bits=(9, 8))
    run("phj", R, S, radix_bits=(8, 9))
    run("phj", R, S, radix_bits=(9, 9))
    run("phj", R, S, radix_bits=(8, 7))
    run("phj", R, S, radix_bits=(6, 5, 5))
    del R, S
if which in ("npj", "all"):
    nr, ns = 1 << 24, 1 << 28
    R = eng.generate(0, nr, nr, 42, 1, datagen.INNER_FACTOR)
    S = eng.generate(1, ns, nr, 42, 2, datagen.OUTER_FACTOR)
    for load in (0.25, 0.5, 0.75, 0.9):
        run("npj", R, S, npj_load=load)
    run("npj", R, S, materialize=False)
    run("npj", R, S, npj_load=0.9, materialize=False)
    run("phj", R, S)                      # PHJ on config 1's shape
    run("phj", R, S, materialize=False)
    del R, S
    nr, ns = 1 << 16, 1 << 30             # config 3
    R = eng.generate(0, nr, nr, 42, 1, datagen.INNER_FACTOR)
    S = eng.generate(1, ns, nr, 42, 2, datagen.OUTER_FACTOR)
    run("npj", R, S, reps=2)
    run("npj", R, S, reps=2, materialize=False)
    run("phj", R, S, reps=2)
    run("phj", R, S, reps=2, materialize=False)
