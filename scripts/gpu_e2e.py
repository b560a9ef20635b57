"""Where the end-to-end time of the pipelined host entry point goes (config 2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
eng = hj.Engine(0)
n = 1 << 27
R = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR); S = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
pin = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(4)]
for p, t in zip(pin, (*R, *S)):
    p.copy_(t)
torch.cuda.synchronize()
h = [p.numpy() for p in pin]
for slices in (os.environ.get("HJB_HOST_SLICE", "default"),):
    for _ in range(2):
        eng.phj((h[0], h[1]), (h[2], h[3]))
    for _ in range(3):
        t0 = time.perf_counter()
        r = eng.phj((h[0], h[1]), (h[2], h[3]))
        t1 = time.perf_counter()
        print(f"slice={slices} wall {1e3*(t1-t0):7.2f} ms  e2e(lib) {r.seconds_e2e*1e3:7.2f}  compute-stream {r.seconds*1e3:7.2f}  h2d {r.phase_ms[5]:6.2f}  d2h {r.phase_ms[6]:6.2f}  build {r.phase_ms[0]:5.2f}")
