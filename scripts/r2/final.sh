#!/bin/bash
# what the driver runs at round end, on N GPUs: tests, smoke, bench (both arms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r02_bench_cpra_${N}gpu.json 2> gpurun_out/r02_bench_cpra_${N}gpu.err; echo "bench rc=$?"
timeout 600 bash scripts/gpu_cli_multi.sh $N > gpurun_out/r02_cli_multi_$N.log 2>&1; echo "cli rc=$?"; tail -6 gpurun_out/r02_cli_multi_$N.log
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench*.json')):
    try:
        l=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f, 'ms', round(l.get('ms_per_step',0),3), 'G/s', round(l['value']/1e9,2), 'e2e', round(l['e2e']['value']/1e9,2), 'roof', (l.get('roofline') or {}).get('frac'), 'step', (l.get('step_roofline') or {}).get('frac_of_hbm_peak'))
    except Exception as e: print(f, e)
PY
