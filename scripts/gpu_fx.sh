#!/bin/bash
# fixed-region scatter (HJB_SCATTER_VARIANT=11) against the default
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=${1:-11}
HJB_SCATTER_VARIANT=$V timeout 600 python -m pytest tests/test_join_gpu.py -m gpu -q --timeout 120 -p no:cacheprovider -x > gpurun_out/tests_fx.log 2>&1; tail -3 gpurun_out/tests_fx.log
for v in $V 3; do
HJB_SCATTER_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v$v.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_v$v.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("variant $v", round(d["ms_per_step"], 3), "ms", d["kernel_ms_per_step"])
PY
done
true
