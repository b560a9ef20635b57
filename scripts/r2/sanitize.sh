#!/bin/bash
# compute-sanitizer over the small parity tests: memcheck (incl. the peer scatter with virtual owners and the skew kernels),
# then racecheck on the shared-memory kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cpra_gpu.py -q -x -p no:cacheprovider -k "virtual_owners or hot_keys or too_small" > gpurun_out/r02_san_cpra.log 2>&1; echo "memcheck cpra rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/r02_san_cpra.log | sort | uniq -c | head
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_join_gpu.py -q -x -p no:cacheprovider -k "pipelined or graph or fingerprint or tiny or special or some_partitions or partition_pass or plans" > gpurun_out/r02_san_join.log 2>&1; echo "memcheck join rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/r02_san_join.log | sort | uniq -c | head
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_join_gpu.py -q -x -p no:cacheprovider -k "tiny or special or partition_pass or plans" > gpurun_out/r02_race_join.log 2>&1; echo "racecheck join rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r02_race_join.log | sort | uniq -c | head
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_cpra_gpu.py -q -x -p no:cacheprovider -k "virtual_owners and (uniform or tiny) or hot_keys" > gpurun_out/r02_race_cpra.log 2>&1; echo "racecheck cpra rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r02_race_cpra.log | sort | uniq -c | head
