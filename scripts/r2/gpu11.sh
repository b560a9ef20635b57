#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
bash scripts/r2/gpu10.sh 8
timeout 1500 python -m pytest tests/test_cpra_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider -k "nccl" > gpurun_out/r2_tests_nccl8.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_tests_nccl8.log
