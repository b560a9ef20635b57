"""Phase clocks of the scatter kernel (HJB_SCATTER_CLOCKS=1): cycles thread 0 of every CTA spent per phase."""
import os, sys, ctypes as C
os.environ["HJB_SCATTER_CLOCKS"] = "1"
os.environ["HJB_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
eng = hj.Engine(0)
n = 1 << 27
R = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR); S = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
out = (C.c_uint64 * 8)()
for _ in range(3):
    eng.phj(R, S)
eng._lib.hjb_debug_counters(eng._ctx, out)      # clear
r = eng.phj(R, S)
eng._lib.hjb_debug_counters(eng._ctx, out)
v = [int(x) for x in out]
tot = sum(v)
names = ["load+rank", "wait A", "plan", "place+carry", "wait C", "stream"]
tiles = 4 * n / 8192
print("join ms", r.seconds * 1e3, "cycles per tile", tot / tiles)
for k, nm in enumerate(names):
    print(f"{nm:12s} {v[k] / tiles:9.0f} cycles/tile {v[k] * 100 / tot:5.1f}%")
