#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests.log
tail -4 gpurun_out/tests.log
timeout 900 python scripts/gpu_variants.py phj > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_partition_join' -s 3 -c 1 -o gpurun_out/prof_join2 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scatter' -s 12 -c 2 -o gpurun_out/prof_scatter3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
