#!/bin/bash
# 2 GPUs: NCCL/IPC parity test, then the CPRA bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_cpra_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "nccl" > gpurun_out/r2_tests_nccl2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_tests_nccl2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2.log 2> gpurun_out/r2_bench_2.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_2.log; tail -5 gpurun_out/r2_bench_2.err
