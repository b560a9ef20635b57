#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1200 python -m pytest tests/test_cpra_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests_cpra.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_tests_cpra.log
