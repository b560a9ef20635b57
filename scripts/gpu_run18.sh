#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; tail -2 gpurun_out/tests.log
timeout 900 python scripts/gpu_configs.py > gpurun_out/configs.log 2>&1; cat gpurun_out/configs.log | cut -c1-260
