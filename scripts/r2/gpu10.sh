#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
HJB_BENCH_CFG5=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$N.log 2> gpurun_out/r2_bench_$N.err; echo "bench rc=$?"; python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_$N.log') if x.startswith('{')][-1])
print('ms', l['ms_per_step'], 'value', l['value']/1e9, 'e2e', l['e2e']['value']/1e9, l['e2e']['ms_per_step'])
print(l['cpra_ms_per_step'], l['kernel_ms_per_step'])
print(l.get('nvlink'))
print(json.dumps(l.get('cpra_cfg5'), indent=1))
print(json.dumps(l.get('cpra_roofline'), indent=1))
PY
tail -5 gpurun_out/r2_bench_$N.err
