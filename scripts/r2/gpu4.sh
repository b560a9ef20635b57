#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/r2/sanity.py > gpurun_out/r2_sanity.log 2>&1; rc=$?; tail -2 gpurun_out/r2_sanity.log
if [ $rc -ne 0 ]; then echo "sanity failed rc=$rc"; tail -30 gpurun_out/r2_sanity.log; exit 1; fi
for shape in 1 2 0; do
HJB_SCATTER_SHAPE=$shape timeout 300 python scripts/r2/sanity.py > gpurun_out/r2_sanity_s$shape.log 2>&1; tail -1 gpurun_out/r2_sanity_s$shape.log
HJB_SCATTER_SHAPE=$shape timeout 300 python scripts/r2/exp.py cfg2,cfg1 phj 1 2>&1 | tee -a gpurun_out/r2_exp4.log
done
HJB_SCATTER_SHAPE=2 HJB_ITEMS=592 timeout 300 python scripts/r2/exp.py cfg2 phj 1 2>&1 | tee -a gpurun_out/r2_exp4.log
HJB_SCATTER_SHAPE=2 HJB_ITEMS=2368 timeout 300 python scripts/r2/exp.py cfg2 phj 1 2>&1 | tee -a gpurun_out/r2_exp4.log
HJB_SCATTER_SHAPE=2 HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scatter_tc' -s 8 -c 1 -o gpurun_out/r2c_prof_scatter_s2 -f python scripts/r2/exp.py cfg2 phj 1 > gpurun_out/r2_ncu_s2.log 2>&1
HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_partition_join' -s 2 -c 1 -o gpurun_out/r2c_prof_join -f python scripts/r2/exp.py cfg2 phj 1 > gpurun_out/r2_ncu_join.log 2>&1
