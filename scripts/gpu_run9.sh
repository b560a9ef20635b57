#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for v in 0 1 2 3 4; do echo "== join variant $v"; HJB_JOIN_VARIANT=$v timeout 600 python scripts/gpu_variants.py phj 2>&1 | head -2; done
