#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory kernels) over the small parity tests,
# including the two-rank fused exchange (TMA bulk copies into the peer's memory)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tests/test_cpra_nccl.py > gpurun_out/san_cpra.log 2>&1; echo "memcheck cpra rc=$?"
grep -E "CPRA_NCCL|ERROR SUMMARY|Invalid|MISMATCH" gpurun_out/san_cpra.log | sort | uniq -c | head
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_join_gpu.py -q -x -p no:cacheprovider -k "pipelined or graph or fingerprint or tiny or special or virtual_owners or some_partitions" > gpurun_out/san_join.log 2>&1; echo "memcheck join rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/san_join.log | sort | uniq -c | head
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_join_gpu.py -q -x -p no:cacheprovider -k "tiny or special or some_partitions or partition_pass" > gpurun_out/race_join.log 2>&1; echo "racecheck join rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/race_join.log | sort | uniq -c | head
