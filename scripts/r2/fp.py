"""which algorithm writes wrong rows?  fingerprint of the rows vs the rows rebuilt from S"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import datagen
eng = hj.Engine(0)
for lr, ls, kind in ((20, 22, 1), (24, 24, 0), (27, 27, 0), (27, 27, 0), (24, 28, 1)):
    R = eng.generate(0, 1 << lr, 1 << lr, 42, 1, datagen.INNER_FACTOR)
    S = eng.generate(kind, 1 << ls, 1 << lr, 42, 2, datagen.OUTER_FACTOR)
    inner = ((S[0].to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF).to(torch.int32)
    torch.cuda.synchronize()                      # the engine runs on its own non-blocking stream
    want = eng.rows_fingerprint(S[0], S[1], inner)
    for a in ("npj", "phj"):
        for rep in range(2):
            r = getattr(eng, a)(R, S)
            k, o, i = r.rows_torch()
            fp = eng.rows_fingerprint(k, o, i)
            eng.synchronize()
            kk = k.to(torch.int64) & 0xFFFFFFFF
            bad_o = int(((kk * datagen.OUTER_FACTOR & 0xFFFFFFFF) != (o.to(torch.int64) & 0xFFFFFFFF)).sum())
            bad_i = int(((kk * datagen.INNER_FACTOR & 0xFFFFFFFF) != (i.to(torch.int64) & 0xFFFFFFFF)).sum())
            print(os.environ.get("HJB_CTA_EMIT"), lr, ls, a, rep, "count", r.count, "fp", "OK" if fp == want else "BAD", "bad rows", bad_o, bad_i, flush=True)
    del R, S
