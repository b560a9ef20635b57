#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
python scripts/gpu_clocks.py 2>&1 | tail -8
timeout 600 python scripts/gpu_variants.py phj 2>&1 | head -2
