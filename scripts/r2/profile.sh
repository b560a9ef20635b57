#!/bin/bash
# round 2 evidence for profiles/: launch list of the bench command + one ncu --set full capture per hot kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
HJB_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_phj.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_launches.log 2>&1
cap() { # name kernel-regex skip count command...
  local name=$1 k=$2 s=$3; shift 3
  HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c 1 -o gpurun_out/r02_prof_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1
}
cap scatter 'k_scatter_tc' 8 python scripts/r2/exp.py cfg2 phj 1
cap join 'k_partition_join' 2 python scripts/r2/exp.py cfg2 phj 1
cap hist 'k_hist_tiles' 8 python scripts/r2/exp.py cfg2 phj 1
cap npj_cfg1 'k_npj_probe' 2 python scripts/r2/exp.py cfg1 npj 1
cap npj_cfg3 'k_npj_probe' 2 python scripts/r2/exp.py cfg3 npj 1
cap join_cfg3 'k_partition_join' 2 python scripts/r2/exp.py cfg3 phj 1
cap bulk 'k_scatter_bulk' 2 python scripts/r2/bulk_one_gpu.py
timeout 300 python scripts/r2/exp.py cfg1,cfg3 npj 1 2>&1 | tee gpurun_out/r02_exp_npj.log
ls -la gpurun_out/r02_* | head -30
