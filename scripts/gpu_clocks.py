import os, sys, ctypes as C
os.environ["HJB_PHASE_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
eng = hj.Engine(0); eng.set_profiling(True)
n = 1 << 27
R = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR); S = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
for _ in range(3):
    r = eng.phj(R, S)
out = (C.c_uint64 * 8)()
eng._lib.hjb_debug_counters(eng._ctx, out)
tot = sum(out)
names = ["task fetch+wait", "bitmap clear", "build1 keys+atomicOr", "rank scan", "build3 load+place", "probe+emit", "-", "-"]
print("join ms", eng.kernel_times()["k_partition_join"])
for k in range(6):
    print(f"{names[k]:24s} {out[k]/1e6:10.1f} Mcycles {100*out[k]/tot:5.1f}%")
