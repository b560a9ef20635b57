#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -12 gpurun_out/tests.log
timeout 900 python scripts/gpu_configs.py > gpurun_out/configs.log 2>&1; cat gpurun_out/configs.log | cut -c1-330
