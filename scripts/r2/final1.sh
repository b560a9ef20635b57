#!/bin/bash
# round 2, one GPU: what the driver runs at round end (tests, smoke, both bench arms), the launch list of the bench command,
# and ncu --set full captures of the kernels of the staged exchange (two virtual owners, 512-way stage A, 12288-tuple join fills)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
HJB_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_phj.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
cap() { local name=$1 k=$2 s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c 1 -o gpurun_out/r02_prof_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
export STAGE_PLAN=9,9,1 STAGE_LOG2=26 HJB_STAGE_COPY=tma
cap stage_scatter512 'k_scatter_tc' 1 python scripts/r2/stage_one_gpu.py
cap stage_copy 'k_peer_copy' 1 python scripts/r2/stage_one_gpu.py
cap stage_join_bigfill 'k_partition_join' 1 python scripts/r2/stage_one_gpu.py
ls -la gpurun_out/r02_prof_stage* 2>/dev/null
