#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/r2/sanity.py > gpurun_out/r2_sanity.log 2>&1; rc=$?; tail -3 gpurun_out/r2_sanity.log
if [ $rc -ne 0 ]; then echo "sanity failed rc=$rc"; tail -30 gpurun_out/r2_sanity.log; exit 1; fi
timeout 1200 python -m pytest tests/test_cpra_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2_tests_cpra.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_tests_cpra.log
for e in 0 1; do
HJB_CTA_EMIT=$e timeout 600 python scripts/r2/exp.py cfg2,cfg1,cfg3 npj,phj 1 2>&1 | tee -a gpurun_out/r2_exp2.log
done
timeout 600 python scripts/r2/exp.py cfg2,cfg1,cfg3 npj,phj 0 2>&1 | tee -a gpurun_out/r2_exp2.log
for mb in 48 64; do
HJB_NPJ_PHASE_MB=$mb timeout 300 python scripts/r2/exp.py cfg1 npj 1 2>&1 | tee -a gpurun_out/r2_exp2.log
done
HJB_SCATTER_SHAPE=1 timeout 300 python scripts/r2/exp.py cfg2 phj 1 2>&1 | tee -a gpurun_out/r2_exp2.log
for shape in 0 1; do
HJB_SCATTER_SHAPE=$shape HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scatter_tc' -s 8 -c 1 -o gpurun_out/r2_prof_scatter_s$shape -f python scripts/r2/exp.py cfg2 phj 1 > gpurun_out/r2_ncu_s$shape.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
