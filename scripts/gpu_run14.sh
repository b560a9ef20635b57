#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for v in 3 7 8 9; do echo "== scatter variant $v"; HJB_SCATTER_VARIANT=$v timeout 600 python scripts/gpu_variants.py phj 2>&1 | head -1; done
HJB_SCATTER_VARIANT=7 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "partition_pass or join_matches_oracle or tiny" 2>&1 | tail -2
HJB_SCATTER_VARIANT=8 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "partition_pass or join_matches_oracle or tiny" 2>&1 | tail -2
