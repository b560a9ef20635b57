#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/r2/sanity.py > gpurun_out/r2_sanity.log 2>&1; rc=$?; tail -2 gpurun_out/r2_sanity.log
if [ $rc -ne 0 ]; then echo "sanity failed rc=$rc"; tail -30 gpurun_out/r2_sanity.log; exit 1; fi
timeout 300 python scripts/r2/exp.py cfg2,cfg1,cfg3 phj 1 2>&1 | tee -a gpurun_out/r2_exp5.log
HJB_JOIN_MINB=4 timeout 300 python scripts/r2/exp.py cfg2,cfg1 phj 1 2>&1 | tee -a gpurun_out/r2_exp5.log
timeout 300 python scripts/r2/exp.py cfg2 phj 0 2>&1 | tee -a gpurun_out/r2_exp5.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_tests.log
