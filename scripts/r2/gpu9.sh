#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_1.log 2> gpurun_out/r2_bench_1.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2_bench_1.log; tail -3 gpurun_out/r2_bench_1.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.log 2> gpurun_out/r2_bench_ref.err ) 2>&1 | grep real; cat gpurun_out/r2_bench_ref.log; tail -3 gpurun_out/r2_bench_ref.err
free -g | head -2; nproc
