"""Quick parity check before anything long runs on the box: small and medium joins of every algorithm against numpy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import hash_join_codes_knl_b200 as hj
from _oracle import numpy_join, sort_rows
eng = hj.Engine(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()
rng = np.random.default_rng(1)
for nr, ns in ((1000, 5000), (100003, 300001), (1 << 21, 1 << 22)):
    rk = rng.permutation(nr * 3)[:nr].astype(np.uint32) + 1
    rv = rk * np.uint32(7)
    sk = rk[rng.integers(0, nr, ns)]
    sk[::7] += np.uint32(3 * nr + 5)
    sv = np.arange(ns, dtype=np.uint32)
    want = numpy_join(rk, rv, sk, sv)
    for algo, opts in (("npj", {}), ("phj", {}), ("phj", {"radix_bits": (8, 8)}), ("phj", {"radix_bits": (3, 5)}), ("phj", {"radix_bits": (11,)})):
        got = getattr(eng, algo)((dev(rk), dev(rv)), (dev(sk), dev(sv)), **opts)
        ok = got.checks() == want.checks() and (sort_rows(*got.rows_numpy()) == want.sorted_rows()).all()
        print(nr, ns, algo, opts, "OK" if ok else "MISMATCH", flush=True)
        assert ok
print("SANITY_OK")
