#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
CT=${2:-16 20}; MO=${3:-staged-serial staged}
for c in $CT; do
for m in $MO; do
HJB_COPY_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --exchange $m --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_sweep_${m}_${c}_$N.json 2> gpurun_out/r02_sweep.err
python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r02_sweep_${m}_${c}_$N.json') if x.startswith('{')][-1])
    print('ctas $c $m', 'ms', round(l['ms_per_step'],3), l.get('cpra_ms_per_step'), l.get('kernel_ms_per_step'))
except Exception as e: print('$m', e)
PY
done
done
