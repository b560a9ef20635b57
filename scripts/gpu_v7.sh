#!/bin/bash
# experiment: local scatter through TMA bulk copies (HJB_SCATTER_VARIANT=7) against the default
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
HJB_SCATTER_VARIANT=9 timeout 600 python -m pytest tests/test_join_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests_v9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_v9.log
tail -4 gpurun_out/tests_v9.log
for v in 9 3; do
HJB_SCATTER_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v$v.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_v$v.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("variant $v", round(d["ms_per_step"], 3), "ms", d["kernel_ms_per_step"])
PY
done
