#!/bin/bash
# the parity suite with the optional fast paths switched off (they are fallbacks of the default paths)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
HJB_GRAPHS=0 HJB_HOST_PIPELINE=0 timeout 600 python -m pytest tests/test_join_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests_nofast.log 2>&1; echo "graphs off, host pipeline off:"; tail -1 gpurun_out/tests_nofast.log
HJB_PHASE_CLOCKS=1 timeout 600 python -m pytest tests/test_join_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "matches_oracle or tiny or special" > gpurun_out/tests_clocks.log 2>&1; echo "phase-clock instantiation:"; tail -1 gpurun_out/tests_clocks.log
