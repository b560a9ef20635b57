#!/bin/bash
# first GPU contact: smoke, the gpu test-suite (all failures, not -x), a sanitizer pass on the small tests, a short bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep -c avx512f /proc/cpuinfo >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
python -c "import os; print(sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0)))" >> gpurun_out/host.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_join_gpu.py -m gpu -q -p no:cacheprovider -k "tiny or special or golden_rows or histogram_kernel" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/tests.log; tail -5 gpurun_out/sanitizer.log; tail -3 gpurun_out/bench.log
