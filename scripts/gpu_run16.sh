#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "npj or join_matches or full_size or special or tiny or duplicate" 2>&1 | tail -2
timeout 600 python scripts/gpu_variants.py npj 2>&1 | head -8
