#!/bin/bash
# multi-GPU check (gpurun --gpus N): NCCL CPRA parity against the oracle, then the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/test_cpra_nccl.py > gpurun_out/cpra_nccl_$N.log 2>&1; echo "rc=$?" >> gpurun_out/cpra_nccl_$N.log
grep -E "cpra world|CPRA_NCCL|rc=|Error|error" gpurun_out/cpra_nccl_$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_$N.log
tail -3 gpurun_out/bench_$N.log | cut -c1-1800
