#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for it in 2048 1184 1024 1480 2960 4096; do
HJB_ITEMS=$it timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_it$it.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_it$it.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("items $it", round(d["ms_per_step"], 3), "ms", d["kernel_ms_per_step"])
PY
done
