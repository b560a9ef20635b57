#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload npj_cfg1 --no-cpu-baseline > gpurun_out/bench_npj.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_npj.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_phj.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scatter|k_partition_join|k_hist$' -s 16 -c 5 -o gpurun_out/prof_phj -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_npj' -s 6 -c 2 -o gpurun_out/prof_npj -f python bench.py --steps 2 --warmup 3 --workload npj_cfg1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_npj.log 2>&1
tail -4 gpurun_out/tests.log; tail -2 gpurun_out/bench.log; tail -2 gpurun_out/bench_npj.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
