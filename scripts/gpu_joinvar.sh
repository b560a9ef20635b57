#!/bin/bash
# join CTA-shape variants after the spill fix
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 0 3 5 6 7; do
HJB_JOIN_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_jv$v.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_jv$v.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("join variant $v", round(d["ms_per_step"], 3), "ms", d["kernel_ms_per_step"]["k_partition_join"])
PY
done
