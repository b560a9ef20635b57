#!/bin/bash
# quick GPU loop: parity tests, then the PHJ and NPJ bench lines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -6 gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload npj_cfg1 --no-cpu-baseline > gpurun_out/bench_npj.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_npj.log
python - <<'PY'
import json
for f in ("gpurun_out/bench.log", "gpurun_out/bench_npj.log"):
    for ln in open(f):
        if ln.startswith("{"):
            d = json.loads(ln)
            print(f, "ms/step", round(d["ms_per_step"], 3), "Gtuples/s", round(d["value"] / 1e9, 2), "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e9, 2))
            print("   kernels", d["kernel_ms_per_step"])
            print("   roofline", d["roofline"])
        elif "rc=" in ln or "Error" in ln or "error" in ln:
            print(f, ln.strip()[:300])
PY
