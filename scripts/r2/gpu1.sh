#!/bin/bash
# round 2, first GPU session: sanity -> test suite -> scatter shapes -> NPJ phase sweep -> configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python scripts/r2/sanity.py > gpurun_out/r2_sanity.log 2>&1; rc=$?; tail -20 gpurun_out/r2_sanity.log
if [ $rc -ne 0 ]; then echo "sanity failed rc=$rc"; exit 1; fi
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_tests.log
for shape in 0 1; do
HJB_SCATTER_SHAPE=$shape timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_shape$shape.log 2>&1; tail -c 1500 gpurun_out/r2_bench_shape$shape.log | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shape', $shape, l['ms_per_step'], l['kernel_ms_per_step'], l['roofline']['frac'])"
done
for mb in 100000 32 48 64; do
HJB_NPJ_PHASE_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --workload npj_cfg1 --no-cpu-baseline --no-e2e > gpurun_out/r2_npj_mb$mb.log 2>&1; python -c "import sys,json; l=json.loads(open('gpurun_out/r2_npj_mb$mb.log').read().strip().splitlines()[-1]); print('npj phase_mb', $mb, l['ms_per_step'], l['kernel_ms_per_step'])"
done
HJB_NPJ_HINTS=0 timeout 300 python bench.py --steps 5 --warmup 3 --workload npj_cfg1 --no-cpu-baseline --no-e2e > gpurun_out/r2_npj_nohints.log 2>&1; python -c "import sys,json; l=json.loads(open('gpurun_out/r2_npj_nohints.log').read().strip().splitlines()[-1]); print('npj nohints', l['ms_per_step'], l['kernel_ms_per_step'])"
timeout 900 python scripts/gpu_configs.py > gpurun_out/r2_configs.log 2>&1; cat gpurun_out/r2_configs.log
