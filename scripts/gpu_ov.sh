#!/bin/bash
# overlapped scatter (default) against the previous kernel (HJB_SCATTER_VARIANT=10)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_join_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x > gpurun_out/tests_ov.log 2>&1; tail -3 gpurun_out/tests_ov.log
for v in 3 10; do
HJB_SCATTER_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v$v.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_v$v.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("variant $v", round(d["ms_per_step"], 3), "ms", d["kernel_ms_per_step"])
PY
done
timeout 120 python scripts/gpu_scatter_clocks.py
