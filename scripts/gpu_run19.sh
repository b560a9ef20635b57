#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scatter' -s 14 -c 1 -o gpurun_out/r01d_prof_scatter -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hist' -s 14 -c 1 -o gpurun_out/r01d_prof_hist -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_npj_probe' -s 3 -c 1 -o gpurun_out/r01d_prof_npj -f python bench.py --steps 2 --warmup 3 --workload npj_cfg1 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/r01d*.ncu-rep
