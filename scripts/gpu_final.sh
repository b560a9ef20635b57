#!/bin/bash
# what the driver runs at round end, plus the ncu evidence pack
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r01}
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --workload npj_cfg1 --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_npj.json 2>/dev/null; echo "npj rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
HJB_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_phj.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench_reference.err
