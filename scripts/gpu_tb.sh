#!/bin/bash
# parity tests + the PHJ bench line (kernel times)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; tail -2 gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.log 2>&1
python - <<'PY'
import json
for ln in open("gpurun_out/bench.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("ms/step", round(d["ms_per_step"], 3), "instr", round(d["ms_per_step_instrumented"], 3), d["kernel_ms_per_step"], d["gpu_launches"])
PY
