#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in 0 111 74 48; do
HJB_PEER_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_2_c$c.log 2> gpurun_out/r2_bench_2_c$c.err; python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_2_c$c.log') if x.startswith('{')][-1])
print('peer_ctas', $c, 'ms', round(l['ms_per_step'],3), l['cpra_ms_per_step'], 'bulk', l['kernel_ms_per_step'].get('k_scatter_bulk'), 'nvlink', l['nvlink']['achieved_gbs_per_direction'])
PY
done
