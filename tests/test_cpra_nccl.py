"""Multi-GPU CPRA over NCCL (one process per GPU).  Run under torchrun on a box with >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/test_cpra_nccl.py
Every rank holds one contiguous chunk of R and S (cpra2.cpp:1724-1731); the global count and
checksums must equal the oracle's CPRA on the whole input, and the union of the ranks' rows must
equal the oracle's rows."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import hash_join_codes_knl_b200 as hj
    from hash_join_codes_knl_b200 import cpra
    from _oracle import oracle_generate, oracle_join, sort_rows
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    eng = hj.Engine(local, use_torch_stream=True)
    fused = cpra.FusedExchange(eng)
    ok = True
    cases = [(200000, 600000, 16), (1 << 20, 1 << 22, 17), (100003, 70001, 18), (150000, 500000, -19),
             (37, 5, 20), (4096, 33, 21), (1, 70000, 22)]          # pieces of a few tuples: unaligned heads and tails only
    rnd = np.random.default_rng(int(os.environ.get("HJB_CPRA_RANDOM_SEED", "1")))
    for i in range(int(os.environ.get("HJB_CPRA_RANDOM", "0"))):   # extra random shapes (same on every rank)
        cases.append((int(rnd.integers(1, 60000)), int(rnd.integers(1, 200000)), 100 + i if i % 3 else -(100 + i)))
    for nr, ns, seed in cases:
        rk, rv, sk, sv, _, _ = oracle_generate(nr, ns, threads=2, seed=abs(seed))
        if seed < 0:
            # skew: a third of the probe side is one heavy-hitter key (all of it lands on one owner), a tenth has no partner
            sk = sk.copy()
            sk[::3] = rk[7]
            sk[5::10] ^= np.uint32(0x10000000)
        want = oracle_join("cpra", rk, rv, sk, sv, threads=4)
        cr = slice(rank * nr // world, (rank + 1) * nr // world)
        cs = slice(rank * ns // world, (rank + 1) * ns // world)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()
        fused.stage_plan = None                                  # the staged plan is made per input shape
        for mode in ("nccl", "fused", "fused", "fused+skew", "staged", "staged", "staged-serial", "staged-parts", "staged-parts"):    # twice: the second step reuses plans / buffers
            if mode == "nccl":
                res = cpra.cpra_join(eng, (dev(rk[cr]), dev(rv[cr])), (dev(sk[cs]), dev(sv[cs])))
            elif mode.startswith("staged"):
                if mode.endswith("parts") and getattr(fused, "stage_parts", 0) != 4:
                    fused.stage_plan = None                      # re-plan: the runs leave in four pieces
                res = cpra.cpra_join_staged(eng, (dev(rk[cr]), dev(rv[cr])), (dev(sk[cs]), dev(sv[cs])), fused,
                                            overlap=not mode.endswith("serial"), parts=4 if mode.endswith("parts") else None)
            else:
                res = cpra.cpra_join_fused(eng, (dev(rk[cr]), dev(rv[cr])), (dev(sk[cs]), dev(sv[cs])), fused,
                                           skew=mode.endswith("skew"))
            got = (res["count"], res["sum_key"], res["sum_outer"], res["sum_inner"])
            rows = list(res["local"].rows_numpy())
            gathered = [None] * world
            dist.all_gather_object(gathered, rows)
            if rank == 0:
                allrows = sort_rows(*(np.concatenate([g[i] for g in gathered]) for i in range(3)))
                good = got == want.checks() and (allrows == want.sorted_rows()).all()
                if mode.endswith("skew") and seed < 0 and ns >= 100000:
                    good = good and res.get("hot_keys", 0) >= 1          # the planted heavy hitter was found
                print(f"cpra {mode} world={world} |R|={nr} |S|={ns} hot={res.get('hot_keys', 0)}: {'OK' if good else 'MISMATCH'} {got} want {want.checks()} "
                      f"split {res['split_ms']:.3f} ms exchange {res['exchange_ms']:.3f} ms join {res['join_ms']:.3f} ms", flush=True)
                ok = ok and good
    dist.destroy_process_group()
    if rank == 0:
        print("CPRA_NCCL_OK" if ok else "CPRA_NCCL_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
