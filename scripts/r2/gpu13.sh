#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/r2/sanity.py > gpurun_out/r2_sanity.log 2>&1; rc=$?; tail -1 gpurun_out/r2_sanity.log
if [ $rc -ne 0 ]; then echo "sanity failed rc=$rc"; tail -30 gpurun_out/r2_sanity.log; exit 1; fi
timeout 300 python scripts/r2/exp.py cfg1,cfg3,cfg2 npj 1 2>&1 | tee -a gpurun_out/r2_exp13.log
HJB_NPJ_CTA_EMIT=1 timeout 300 python scripts/r2/exp.py cfg3 npj 1 2>&1 | tee -a gpurun_out/r2_exp13.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_tests.log
