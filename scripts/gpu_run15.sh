#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -12 gpurun_out/tests.log
for v in 3 8; do echo "== scatter variant $v"; HJB_SCATTER_VARIANT=$v timeout 600 python scripts/gpu_variants.py phj 2>&1 | head -7; done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_join_gpu.py -m gpu -q -p no:cacheprovider -k "partition_pass or tiny or special" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitizer.log
