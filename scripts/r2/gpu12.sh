#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpra_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "virtual or hot or tiles" 2>&1 | tail -3
for b in 2 3; do
HJB_BULK_BUFFERS=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_2_b$b.log 2> gpurun_out/r2_bench_2_b$b.err; python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/r2_bench_2_b$b.log') if x.startswith('{')][-1])
print('buffers', $b, 'ms', round(l['ms_per_step'],3), l['cpra_ms_per_step'], 'bulk', l['kernel_ms_per_step'].get('k_scatter_bulk'), 'nvlink', l['nvlink']['achieved_gbs_per_direction'])
PY
done
