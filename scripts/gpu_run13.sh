#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
for v in 3 5 6; do echo "== scatter variant $v"; HJB_SCATTER_VARIANT=$v timeout 600 python scripts/gpu_variants.py phj 2>&1 | head -1; done
HJB_SCATTER_VARIANT=6 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "partition_pass or join_matches_oracle" 2>&1 | tail -2
