#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python scripts/gpu_e2e.py
HJB_HOST_SLICE=2097152 python scripts/gpu_e2e.py
HJB_HOST_PIPELINE=0 python scripts/gpu_e2e.py
