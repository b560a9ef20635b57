#!/usr/bin/env python
"""bench.py -- join throughput, (|R|+|S|) tuples/s, on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

N = 1 (default): BASELINE.json config 2 -- PHJ, |R| = |S| = 2^27 unique 32-bit keys (S a
permutation of R's key set), 32-bit payloads, materialised 3-column result.  The same run also
measures configs 1 and 3 (NPJ and PHJ each) and the one-GPU base of the weak-scaling curve; they go
into the line as `configs` / `weak_scaling_base`.
N > 1 (under torchrun, one rank per GPU): the CPRA path, weak scaling at 2^28 + 2^28 tuples
per GPU, so that N = 8 is exactly config 4 (2^31 x 2^31): count by owner -> all-gather of the count
matrix -> GPU-assign scatter straight into the owners' buffers over NVLink -> local PHJ -> all-reduce
of the checksums, all enqueued on one stream.

A step = one whole join of one batch.  `value` has the inputs resident in HBM (the reference's
timed region, npj.cpp:861-918); `e2e` goes through the host entry points of the C ABI with pinned host
buffers (H2D of the inputs and D2H of the rows inside the timed region).  Every step's
count / checksums are checked against the analytic expectation (count = |S|, checksums =
column sums of S), a wrong step aborts the run.  One JSON line on stdout (rank 0).

--impl reference: the reference's own CPU program (oracle/_ref, compiled from /root/reference where it
lies) on the box's host cores; nothing of this repository's library is loaded in that arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "join throughput (R+S tuples/sec)"
UNIT = "tuples/s"
MASK64 = (1 << 64) - 1
NVLINK_REF_GBS = 770.0          # measured peer-copy bandwidth per direction (B200_PROFILING.md); nominal 900


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons while the timed regions run (B200_PROFILING.md): NVML polled every
    few milliseconds (the timed region is tens of milliseconds long), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self._halt = gpu_index, threading.Event()
        self.sm, self.max_sm, self.reasons, self.source = [], 0.0, set(), None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        self.source = "nvml"
        while not self._halt.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            self._halt.wait(0.004)

    def _run_smi(self):
        self.source = "nvidia-smi"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                s = [x.strip() for x in out.split(",")]
                self.sm.append(float(s[1]))
                self.max_sm = max(self.max_sm, float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm or None,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": self.source}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ----------------------------------------------------------------------------- reference / CPU arm
# Nothing below imports hash_join_codes_knl_b200: the relation files are written with numpy.

REF_INNER_FACTOR, REF_OUTER_FACTOR = 0x6587F97D, 0xDF56B8FB


def host_threads():
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except Exception:
        cpus = list(range(os.cpu_count() or 1))
    return cpus


def has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable"):
                    return int(ln.split()[1]) / 2**20
    except OSError:
        pass
    return 0.0


def host_relations(nr, ns, foreign_keys, seed=42):
    """Relations of the workload's shape, numpy only (write.cpp's semantics: distinct non-zero keys,
    payload = key * odd factor, S either a permutation of R's key set or foreign keys -- every key once,
    the rest uniform picks -- and both shuffled)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    keys = (np.arange(1, nr + 1, dtype=np.uint32) * np.uint32(0x9E3779B1))          # odd multiplier: distinct, non-zero
    rk = keys[rng.permutation(nr)]
    if foreign_keys:
        sk = np.concatenate([keys, keys[rng.integers(0, nr, ns - nr)]]) if ns > nr else keys[:ns].copy()
        sk = sk[rng.permutation(ns)]
    else:
        sk = keys[rng.permutation(nr)][:ns]
    with np.errstate(over="ignore"):
        rv, sv = rk * np.uint32(REF_INNER_FACTOR), sk * np.uint32(REF_OUTER_FACTOR)
    return rk, rv, sk, sv


def write_relation_files(d, rk, rv, sk, sv):
    """the four raw files the reference reads from its CWD (write.cpp:1824-1865, npj.cpp:1013-1039)"""
    rk.tofile(os.path.join(d, f"ik_{rk.size}.txt"))
    rv.tofile(os.path.join(d, f"iv_{rk.size}.txt"))
    sk.tofile(os.path.join(d, f"ok_{sk.size}.txt"))
    sv.tofile(os.path.join(d, f"ov_{sk.size}.txt"))


def run_reference_program(prog, threads, d, nr, ns, timeout_s):
    """One run of the UNMODIFIED reference program (oracle/_ref/<prog>, compiled from /root/reference by
    oracle/Makefile) in directory d; returns the seconds it prints (its own timer)."""
    exe = os.path.join(ROOT, "oracle", "_ref", prog)
    args = [exe, str(threads), str(ns), str(nr)]
    if prog != "cpra":
        args.append("1")
    out = subprocess.run(args, cwd=d, capture_output=True, text=True, timeout=timeout_s)
    if out.returncode != 0:
        raise RuntimeError(f"{prog} exited {out.returncode}: {out.stderr[-300:]}")
    lines = [ln for ln in out.stdout.strip().splitlines() if ln and not ln.startswith("copy")]
    return float(lines[-1].split()[0])


def cpu_join(log2_r, log2_s, steps, warmup, npj):
    """The reference's CPU implementation, all usable host threads.  Preferred: the reference's own program
    (kind "reference"); else the oracle port.  PHJ's shipped program performs no join (its join phase is
    commented out, phj.cpp:1869-1924), so the partitioned path is timed with the reference's complete
    partitioned join, cpra (cpra2.cpp)."""
    nr, ns = 1 << log2_r, 1 << log2_s
    prog = "npj" if npj else "cpra"
    rk, rv, sk, sv = host_relations(nr, ns, foreign_keys=npj)
    cpus = host_threads()
    use_ref = (os.path.exists(os.path.join(ROOT, "oracle", "_ref", prog)) and has_avx512()
               and cpus == list(range(len(cpus))))        # the reference pins thread t to CPU t (makefile -DSCATTER)
    threads = min(len(cpus), 256)                          # repo_offset[256], cpra2.cpp:1861
    secs, kind_used, note = [], None, ""
    if use_ref:
        try:
            with tempfile.TemporaryDirectory(prefix="hjref_") as d:
                write_relation_files(d, rk, rv, sk, sv)
                for i in range(warmup + steps):
                    s = run_reference_program(prog, threads, d, nr, ns, timeout_s=900)
                    if i >= warmup:
                        secs.append(s)
            kind_used = "reference"
            note = f"oracle/_ref/{prog} (unmodified {('cpra2' if prog == 'cpra' else prog)}.cpp, g++ -O3 -march=skylake-avx512)"
        except Exception as e:            # never let the baseline leg take the bench down
            note = f"reference program failed ({type(e).__name__}: {e}); "
            secs = []
    if not secs:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle
        _oracle.build_oracle()
        threads = min(len(cpus), 64)
        for i in range(warmup + steps):
            r = _oracle.oracle_join(prog, rk, rv, sk, sv, threads=threads, materialize=False)
            assert r.count == ns
            if i >= warmup:
                secs.append(r.seconds)
        kind_used = "port"
        note += "oracle/hj_oracle.c scalar restatement"
    sec = sum(secs) / len(secs)
    return {"value": (nr + ns) / sec, "unit": UNIT, "cores": threads, "kind": kind_used,
            "sample": f"{prog.upper()} |R|=2^{log2_r} x |S|=2^{log2_s} ({note}), mean of {len(secs)} run(s), {sec:.3f} s each",
            "seconds": sec}


def workload_label(workload, world, log2_per_gpu):
    """config.workload, identical in both arms"""
    return {"phj_cfg2": "PHJ 2^27 x 2^27 unique keys (BASELINE config 2)",
            "npj_cfg1": "NPJ 2^24 x 2^28 foreign keys (BASELINE config 1)",
            "cpra_cfg4": f"CPRA 2^{log2_per_gpu} + 2^{log2_per_gpu} tuples per GPU x {world} GPUs "
                         f"(N=8 is BASELINE config 4, 2^31 x 2^31)"}[workload]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = args.workload if args.workload != "auto" else ("phj_cfg2" if args.gpus == 1 else "cpra_cfg4")
    t0 = time.time()
    roomy = mem_available_gb() >= 24.0
    if workload == "npj_cfg1":
        shape = (24, 28) if roomy else (22, 26)
        full = roomy
    else:
        # config 2 at full size when the host has the memory; config 4 (2^31 x 2^31) does not fit a CPU run:
        # the largest power of two that does
        shape = (27, 27) if roomy else (25, 25)
        full = roomy and workload == "phj_cfg2"
    steps = max(1, min(args.steps, 6))              # a run is seconds long: the whole arm ends within minutes
    base = cpu_join(shape[0], shape[1], steps, min(args.warmup, 1), npj=workload == "npj_cfg1")
    note = ("CPU reference, throughput in (R+S) tuples/s does not depend on the GPU count. "
            + ("Full size of the config. " if full else "Bounded sample of the workload. ")
            + ("" if workload == "npj_cfg1" else "The reference's cpra (cpra2.cpp) stands in for PHJ: the shipped phj.cpp "
               "performs no join (phj.cpp:1869-1924 commented out); cpra is its complete partitioned join. "))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["seconds"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_label(workload, args.gpus, args.log2_per_gpu or 28),
                       "sample": base["sample"], "full_size": full, "note": note},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- our arm

def expected_checks(eng, sk, sv, ns):
    import torch
    from hash_join_codes_knl_b200 import datagen
    inner = (sk.to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF
    return (ns, eng.column_sum(sk), eng.column_sum(sv), int(inner.sum().item()) & MASK64)


def step_bytes(algo, nr, ns, matches, passes=2, npj_load=0.75):
    """algorithmic bytes of one whole join"""
    n = nr + ns
    if algo == "npj":
        table = 8 * nr / npj_load
        return 8 * n + table + 8 * nr + (8 * ns if table > (64 << 20) else 0) + 12 * matches
    return n * (20 * passes + 8) + 12 * matches


def kernel_bytes(nr, ns, matches):
    """algorithmic bytes per LAUNCH SET of a kernel in one step, given its launches L"""
    n = nr + ns
    return {
        "k_hist": lambda L: 4 * n * (L / 2),                   # 4 B/tuple; R and S launches alternate
        "k_scatter": lambda L: 16 * n * (L / 2),               # 8 B read + 8 B written per tuple
        "k_scatter_bulk": lambda L: 16 * n * (L / 2),          # the same bytes, 8 of them over NVLink
        "k_partition_join": lambda L: 8 * n + 12 * matches,    # read both partitioned relations, write 12 B/match
        "k_npj_probe": lambda L: 8 * ns + 12 * matches + (0 if nr * 16 <= (64 << 20) else 8 * ns),
        "k_npj_build": lambda L: 8 * nr + 8 * (4 * nr / 3) + 8 * nr,    # read R, init the table (load 0.75), write slots
    }


def measure_join(eng, algo, R, S, want, steps, warmup, peak, traffic=None):
    """steps joins of device-resident columns, every one verified; plus one instrumented step for the
    per-kernel times.  Returns the compact record that goes into `configs`."""
    import torch
    nr, ns = R[0].numel(), S[0].numel()
    eng.set_profiling(False)
    for _ in range(warmup):
        r = getattr(eng, algo)(R, S)
        assert r.checks() == want, f"{algo}: wrong result {r.checks()} != {want}"
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    launches = 0
    for _ in range(steps):
        r = getattr(eng, algo)(R, S)
        launches += r.kernel_launches
        if r.checks() != want:
            raise SystemExit(f"{algo}: timed step produced a wrong result")
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    eng.set_profiling(True)
    getattr(eng, algo)(R, S)
    r = getattr(eng, algo)(R, S)
    kt = {k: (v[0], v[1]) for k, v in eng.kernel_times().items() if v[1]}
    eng.set_profiling(False)
    passes = max(1, round(kt.get("k_scatter", (0, 4))[1] / 2)) if algo == "phj" else 0
    b = step_bytes(algo, nr, ns, want[0], passes)
    kb = kernel_bytes(nr, ns, want[0])
    dom = max(kt.items(), key=lambda kv: kv[1][0])[0] if kt else None
    rec = {"algorithm": algo, "inner_tuples": nr, "outer_tuples": ns, "ms_per_step": round(ms, 4),
           "tuples_per_s": (nr + ns) / (ms * 1e-3), "verified": True, "radix_passes": passes,
           "algorithmic_bytes": b, "step_frac_of_hbm_peak": round(b / (ms * 1e-3) / 1e9 / peak, 4),
           "kernel_ms": {k: round(v[0], 4) for k, v in sorted(kt.items())}, "gpu_launches_per_step": launches // steps}
    if dom in kb:
        per_launch = kb[dom](kt[dom][1]) / kt[dom][1]
        rec["dominant"] = {"kernel": dom, "frac": round(per_launch / (kt[dom][0] / kt[dom][1] * 1e-3) / 1e9 / peak, 4),
                           "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": round(kt[dom][0] / kt[dom][1], 4),
                           "traffic": (traffic or {}).get(dom)}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "phj_cfg2", "npj_cfg1", "cpra_cfg4"])
    ap.add_argument("--log2-per-gpu", type=int, default=0, help="override tuples per relation per GPU (2^k)")
    ap.add_argument("--exchange", default="staged", choices=["staged", "staged-serial", "fused", "nccl"],
                    help="N>1: staged (local pass by owner and sub-partition, TMA copies of whole runs beside the passes; "
                         "-serial: copies on the main stream), fused (GPU-assign pass storing straight into the owners' buffers), "
                         "or split + NCCL all-to-all")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="N=1: skip the configs 1 / 3 / weak-scaling-base table")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import hash_join_codes_knl_b200 as hj
    from hash_join_codes_knl_b200 import cpra as cpra_mod, datagen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run, one rank per GPU")
        args.gpus = world
    torch.cuda.set_device(local)
    devname = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")      # the small collectives between the passes must not queue behind them
        dist.init_process_group("nccl", device_id=torch.device(devname))
    workload = args.workload if args.workload != "auto" else ("phj_cfg2" if world == 1 else "cpra_cfg4")

    eng = hj.Engine(local, use_torch_stream=True)
    fused = cpra_mod.FusedExchange(eng) if (world > 1 and args.exchange != "nccl") else None

    def cpra_step(inner, outer):
        if fused is not None and args.exchange.startswith("staged") and hasattr(inner[0], "is_cuda"):
            return cpra_mod.cpra_join_staged(eng, inner, outer, fused, overlap=args.exchange == "staged")
        if fused is not None:
            return cpra_mod.cpra_join_fused(eng, inner, outer, fused)       # also the host-memory step of every exchange mode
        return cpra_mod.cpra_join(eng, inner, outer)

    # ---- synthetic inputs, generated on the device (identical to datagen's numpy mirror)
    if workload == "cpra_cfg4":
        k = args.log2_per_gpu or 28
        nr_g = ns_g = 1 << k
        nr_tot, ns_tot = nr_g * world, ns_g * world
        rk, rv = eng.generate(0, nr_g, nr_tot, 42, 1, datagen.INNER_FACTOR, first=rank * nr_g, total=nr_tot)
        sk, sv = eng.generate(0, ns_g, nr_tot, 42, 2, datagen.OUTER_FACTOR, first=rank * ns_g, total=ns_tot)
        algo = "cpra"
    else:
        nr_tot, ns_tot, kind = datagen.workload(workload)
        if args.log2_per_gpu:
            scale = (1 << args.log2_per_gpu) / ns_tot
            nr_tot, ns_tot = max(1, int(nr_tot * scale)), 1 << args.log2_per_gpu
        nr_g, ns_g = nr_tot, ns_tot
        rk, rv = eng.generate(0, nr_tot, nr_tot, 42, 1, datagen.INNER_FACTOR)
        sk, sv = eng.generate(kind, ns_tot, nr_tot, 42, 2, datagen.OUTER_FACTOR)
        algo = "npj" if workload == "npj_cfg1" else "phj"
    want_local = expected_checks(eng, sk, sv, ns_g)      # every probe tuple has exactly one partner
    if world > 1:
        want = cpra_mod.reduce_checks(*want_local, torch.device(devname))
    else:
        want = want_local

    def step_device():
        if algo == "cpra":
            r = cpra_step((rk, rv), (sk, sv))
            got = (r["count"], r["sum_key"], r["sum_outer"], r["sum_inner"])
            return got, r["local"].kernel_launches, r
        r = getattr(eng, algo)((rk, rv), (sk, sv))
        return r.checks(), r.kernel_launches, r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the workspace so the timed region never allocates)
    eng.set_profiling(False)
    for _ in range(args.warmup):
        got, _, r_warm = step_device()
        assert got == want, f"warm-up step produced a wrong result: {got} != {want}"
    # full-row check once, on the device: the result rows as a multiset (order-independent 128-bit
    # fingerprint, hjb_rows_fingerprint) against the rows rebuilt from S -- (key, outer payload,
    # key * INNER_FACTOR) for every probe tuple; summed / xor-ed over the ranks for CPRA
    res_warm = r_warm["local"] if algo == "cpra" else r_warm
    fp_got = eng.rows_fingerprint(*res_warm.rows_torch())
    inner_col = ((sk.to(torch.int64) & 0xFFFFFFFF) * datagen.INNER_FACTOR & 0xFFFFFFFF).to(torch.int32)
    fp_want = eng.rows_fingerprint(sk, sv, inner_col)
    del inner_col
    if world > 1:
        def to_i64(x):
            return x - (1 << 64) if x >= (1 << 63) else x
        mine = torch.tensor([to_i64(v) for v in (*fp_got, *fp_want)], dtype=torch.int64, device=devname)
        allv = torch.empty(4 * world, dtype=torch.int64, device=devname)
        dist.all_gather_into_tensor(allv, mine)               # NCCL has no xor reduction: gather, fold on the host
        rows4 = [[int(x) & MASK64 for x in row] for row in allv.view(world, 4).cpu().tolist()]
        fold = [0, 0, 0, 0]
        for row in rows4:
            fold = [(fold[0] + row[0]) & MASK64, fold[1] ^ row[1], (fold[2] + row[2]) & MASK64, fold[3] ^ row[3]]
        fp_got, fp_want = (fold[0], fold[1]), (fold[2], fold[3])
    if fp_got != fp_want:
        raise SystemExit(f"result rows differ from the expected multiset: {fp_got} != {fp_want}")
    del res_warm, r_warm
    # ---- timed: device-resident.  Two timed regions of `steps` steps each: the first as a user runs it
    # (headline `value`), the second with a CUDA event pair around every kernel (per-kernel times, roofline).
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    ktimes = {}
    phases = np.zeros(8)
    extra = {"split_ms": 0.0, "exchange_ms": 0.0, "join_ms": 0.0, "step_ms": 0.0, "copy_r_done_ms": 0.0, "copy_s_done_ms": 0.0}
    recv_stats = []
    stage_plan = {}

    def timed_region(instrumented):
        eng.set_profiling(instrumented)
        got, _, _ = step_device()                  # switching the event pairs on or off re-plans the launch path
        assert got == want
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_launch = 0
        barrier()
        ev0.record()
        for _ in range(args.steps):
            got, nl, r = step_device()
            n_launch += nl
            if got != want:
                raise SystemExit(f"timed step produced a wrong result: {got} != {want}")
            if not instrumented:
                if algo == "cpra":
                    for key in extra:
                        extra[key] += r.get(key, 0.0)
                    recv_stats.append(r["recv_tuples"])
                    if r.get("stage_plan"):
                        stage_plan.update(r["stage_plan"])
                continue
            for name, (ms, n) in eng.kernel_times().items():
                a = ktimes.setdefault(name, [0.0, 0])
                a[0] += ms
                a[1] += n
            phases[:] += np.array(r["local"].phase_ms if algo == "cpra" else r.phase_ms)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=devname)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n_launch

    ms_total, launches = timed_region(False)
    ms_instrumented, launches_instrumented = timed_region(True)
    eng.set_profiling(False)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = (nr_tot + ns_tot) / (ms_step * 1e-3)

    # ---- timed: end to end through the host entry points of the C ABI, pinned host buffers
    e2e = None
    if not args.no_e2e:
        pin = [torch.empty(t.numel(), dtype=torch.int32).pin_memory() for t in (rk, rv, sk, sv)]
        for p, t in zip(pin, (rk, rv, sk, sv)):
            p.copy_(t)
        torch.cuda.synchronize()
        hrk, hrv, hsk, hsv = (p.numpy() for p in pin)
        e2e_steps = max(2, min(args.steps, 5))

        def step_e2e():
            if algo == "cpra":
                r = cpra_step((hrk, hrv), (hsk, hsv))           # hjb_cpra_count_async_host ... hjb_cpra_finish_host
                return (r["count"], r["sum_key"], r["sum_outer"], r["sum_inner"]), r["local"].count
            r = getattr(eng, algo)((hrk, hrv), (hsk, hsv))
            return r.checks(), r.count
        got, _ = step_e2e()
        assert got == want
        barrier()
        t0 = time.perf_counter()
        rows_out = 0
        for _ in range(e2e_steps):
            got, rows_out = step_e2e()
            if got != want:
                raise SystemExit("e2e step produced a wrong result")
        barrier()
        sec = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([sec], device=devname, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        e2e = {"value": (nr_tot + ns_tot) / sec, "unit": UNIT, "h2d_bytes_per_step": 8 * (nr_g + ns_g) * world,
               "d2h_bytes_per_step": 12 * (ns_tot if world > 1 else rows_out), "ms_per_step": sec * 1e3, "steps": e2e_steps,
               "path": ("Engine.%s(host columns) -> hjb_%s_host" % (algo, algo)) if algo != "cpra" else
                       "cpra_join_fused(host columns) -> hjb_cpra_count_async_host / hjb_cpra_scatter_async / "
                       "hjb_cpra_join_async / hjb_cpra_finish_host (every rank: its chunk in, its share of the rows out)"}
        del pin, hrk, hrv, hsk, hsv

    # ---- N > 1: BASELINE config 5 (|R| = 2^27, |S| = 2^30, Zipf theta = 1 probe keys, 50 % of the probe tuples match),
    # sharded like config 4, with and without the heavy-hitter path; every rank takes part, rank 0 reports
    cfg5 = None
    if world > 1 and algo == "cpra" and (world == 8 or os.environ.get("HJB_BENCH_CFG5")) and not args.log2_per_gpu:
        try:
            del rk, rv, sk, sv
            torch.cuda.empty_cache()
            nr5, ns5 = (1 << 27) // world, (1 << 30) // world
            r5 = eng.generate(0, nr5, 1 << 27, 42, 1, datagen.INNER_FACTOR, first=rank * nr5, total=1 << 27)
            s5 = eng.generate(2, ns5, 1 << 27, 42, 2, datagen.OUTER_FACTOR, first=rank * ns5, total=1 << 30, theta=1.0, selectivity=0.5)
            cfg5 = {"workload": f"config 5: 2^27 x 2^30, Zipf theta = 1 probe keys, 50 % selectivity, {world} GPUs"}
            checks = {}
            for mode in ("static_ownership", "hot_keys_local"):
                skew = mode == "hot_keys_local"
                for _ in range(2):
                    r = cpra_mod.cpra_join_fused(eng, r5, s5, fused, skew=skew)
                barrier()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                n5 = 4
                join_ms = 0.0
                for _ in range(n5):
                    r = cpra_mod.cpra_join_fused(eng, r5, s5, fused, skew=skew)
                    join_ms += r["join_ms"]
                ev1.record()
                barrier()
                stats = torch.tensor([ev0.elapsed_time(ev1) / n5, join_ms / n5, float(r["recv_tuples"][1]), float(r["hot_outer_tuples"])],
                                     dtype=torch.float64, device=devname)
                allst = torch.empty(world * 4, dtype=torch.float64, device=devname)
                dist.all_gather_into_tensor(allst, stats)
                st = allst.view(world, 4).cpu()
                checks[mode] = (r["count"], r["sum_key"], r["sum_outer"], r["sum_inner"])
                ms5 = float(st[:, 0].max())
                cfg5[mode] = {"ms_per_step": round(ms5, 4), "tuples_per_s": ((1 << 27) + (1 << 30)) / (ms5 * 1e-3),
                              "recv_probe_tuples_max_over_mean": round(float(st[:, 2].max() / st[:, 2].mean()), 4),
                              "local_join_ms_max_over_mean": round(float(st[:, 1].max() / st[:, 1].mean()), 4),
                              "hot_keys": r["hot_keys"], "hot_probe_tuples_kept_local": int(st[:, 3].sum()), "count": r["count"]}
            cfg5["verified"] = (checks["static_ownership"] == checks["hot_keys_local"]
                                and abs(checks["static_ownership"][0] / (1 << 30) - 0.5) < 0.01)
            cfg5["check"] = "both modes give the same count and checksums; count / |S| within 0.5 +- 0.01"
            del r5, s5
        except Exception as e:
            cfg5 = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel, from the per-launch CUDA events of the timed steps
    peak, peak_src = measured_peak_gbs()
    n_in = nr_g + ns_g                              # tuples this GPU partitions per pass / joins
    alg = kernel_bytes(nr_g, ns_g, ns_g)
    staged = algo == "cpra" and world > 1 and args.exchange.startswith("staged")
    if staged:
        # two passes over this GPU's tuples whatever the number of launches (the local pass and the join run once per part)
        alg["k_hist"] = lambda L: 4 * n_in * 2
        alg["k_scatter"] = lambda L: 16 * n_in * 2
    per_step = {k: (v[0] / args.steps, v[1] / args.steps) for k, v in ktimes.items() if v[1]}
    dom = max(per_step.items(), key=lambda kv: kv[1][0])[0] if per_step else None
    ncu_traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ncu_traffic = json.load(f)
    except Exception:
        pass
    roof = None
    if dom in alg:
        ms_k, launches_k = per_step[dom]
        per_launch_bytes = alg[dom](launches_k) / launches_k
        per_launch_ms = ms_k / launches_k
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
        tr = ncu_traffic.get(workload, {}).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": tr, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": per_launch_ms,
                "launches_per_step": launches_k, "share_of_step": ms_k / (ms_instrumented / args.steps),
                "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                  "ncu --set full)" if tr else None}
    passes = 2 if staged else max(1, round(per_step.get("k_scatter", (0, 4))[1] / 2))     # local scatter launches come in (R, S) pairs
    sb = step_bytes("npj" if algo == "npj" else "phj", nr_g, ns_g, ns_g, passes)
    if staged:
        sb += 16 * n_in * (world - 1) / world                # the copies: a run for another owner is read from the staging columns and written into that owner's columns
    elif algo == "cpra" and world > 1:
        sb += 20 * n_in + 16 * n_in * (world - 1) / world    # GPU-assign pass (hist + scatter) and the receive-side writes / send-side reads
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "ms_per_step_instrumented": ms_instrumented / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": workload_label(workload, world, nr_g.bit_length() - 1),
                   "inner_tuples": nr_tot, "outer_tuples": ns_tot, "materialize": True, "algorithm": algo,
                   "exchange": (args.exchange if world > 1 else None),
                   "exchange_plan": (dict(stage_plan, copies="k_peer_copy (TMA)" if os.environ.get("HJB_STAGE_COPY") == "tma" else "copy engines")
                                     if stage_plan else None),
                   "l2_policy": "inputs (%.1f GiB per GPU) and every intermediate exceed the 126 MB L2; no flush needed"
                                % (8 * (nr_g + ns_g) / 2**30),
                   "result_check": "count and 3 checksums verified every step; all rows verified once as a multiset (device fingerprint)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
        "kernel_ms_per_step": {k: round(v[0], 4) for k, v in sorted(per_step.items())},
        "phase_ms_per_step": [round(float(x) / args.steps, 4) for x in phases],
        "step_roofline": {"algorithmic_bytes_per_gpu": sb, "achieved_gbs_per_gpu": sb / (ms_step * 1e-3) / 1e9,
                          "frac_of_hbm_peak": sb / (ms_step * 1e-3) / 1e9 / peak, "radix_passes_local": passes},
    }
    if algo == "cpra":
        line["cpra_ms_per_step"] = {k: round(v / args.steps, 4) for k, v in extra.items()}
        if recv_stats:
            line["recv_tuples_rank0"] = [int(x) for x in recv_stats[-1]]
        if world > 1:
            sent = 8 * n_in * (world - 1) / world                # bytes each GPU stores into its peers per step
            cm = line["cpra_ms_per_step"]
            bulk_ms = per_step.get("k_scatter_bulk", (0.0, 0))[0]
            if staged:
                # the copies run on a side stream beside the passes: from "R is staged" to "every rank's copies of S are through"
                bulk_ms = cm["exchange_ms"]
            line["nvlink"] = {"bytes_out_per_gpu": sent, "scatter_kernel_ms": round(bulk_ms, 4),
                              "achieved_gbs_per_direction": sent / (bulk_ms * 1e-3) / 1e9 if bulk_ms else None,
                              "reference_gbs": NVLINK_REF_GBS,
                              "note": ("copies (copy engines, or k_peer_copy with HJB_STAGE_COPY=tma) on the side stream: span from the first relation staged to the last copy through on every rank "
                                       "(stage A of S, the local passes and the joins of the pieces that have arrived run beside it)" if staged else "k_scatter_bulk's event-timed launches (R and S)")
                                      + "; measured peer-copy bandwidth per direction 770 GB/s (B200_PROFILING.md), nominal 900"}
            # SURVEY 8d's serial model for CPRA at G GPUs: HBM bytes / HBM bandwidth + NVLink bytes / NVLink bandwidth.
            # `as_built`: the bytes this code moves (GPU-assign + `passes` local passes + join); `survey`: the survey's
            # two-pass figure (GPU-assign + ONE local pass + join) that BASELINE's 12.5 ms target is derived from.
            rl = {"nvlink_bytes_out_per_gpu": sent}
            for tag, hbm_b in (("as_built", sb), ("survey", n_in * 48 + 16 * n_in * (world - 1) / world + 12 * ns_g)):
                for ptag, bw_h, bw_n in (("measured_peaks", peak, NVLINK_REF_GBS), ("nominal_peaks", 8000.0, 900.0)):
                    t_ms = (hbm_b / bw_h + sent / bw_n) / 1e6
                    rl.setdefault(tag, {"hbm_bytes_per_gpu": hbm_b})[ptag] = {"hbm_gbs": bw_h, "nvlink_gbs": bw_n, "serial_model_ms": t_ms,
                                                                              "frac": t_ms / ms_step}
            line["cpra_roofline"] = rl
        if cfg5 is not None:
            line["cpra_cfg5"] = cfg5
    # ---- N = 1: the other single-GPU configs of BASELINE.json, measured in the same run
    if world == 1 and not args.no_configs and workload == "phj_cfg2" and not args.log2_per_gpu:
        del rk, rv, sk, sv
        torch.cuda.empty_cache()
        cfgs = {}
        small = max(3, min(args.steps, 5))
        for name, label in (("npj_cfg1", "config 1: 2^24 x 2^28 foreign keys"), ("small_cfg3", "config 3: 2^16 x 2^30 foreign keys")):
            try:
                nr, ns, kind = datagen.workload(name)
                R = eng.generate(0, nr, nr, 42, 1, datagen.INNER_FACTOR)
                S = eng.generate(kind, ns, nr, 42, 2, datagen.OUTER_FACTOR)
                w = expected_checks(eng, S[0], S[1], ns)
                cfgs[label] = [measure_join(eng, a, R, S, w, small, 2, peak, ncu_traffic.get(name + "_" + a)) for a in ("npj", "phj")]
                del R, S
                torch.cuda.empty_cache()
            except Exception as e:
                cfgs[label] = {"error": f"{type(e).__name__}: {e}"}
        line["configs"] = cfgs
        try:
            n = 1 << 28
            R = eng.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
            S = eng.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
            w = expected_checks(eng, S[0], S[1], n)
            rec = measure_join(eng, "phj", R, S, w, small, 2, peak)
            rec["note"] = ("what ONE GPU does per step of the weak-scaling run (2^28 + 2^28 tuples) without an exchange: "
                           "the like-for-like base of the N = 2, 4, 8 CPRA values")
            line["weak_scaling_base"] = rec
            del R, S
        except Exception as e:
            line["weak_scaling_base"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_cpu_baseline:
        try:
            base = cpu_join(25 if workload != "npj_cfg1" else 22, 25 if workload != "npj_cfg1" else 26, 2, 1, npj=workload == "npj_cfg1")
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
