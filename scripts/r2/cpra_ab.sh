#!/bin/bash
# N GPUs: the NCCL / IPC parity script, then the CPRA bench line under every exchange mode (device-resident part only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
MODES=${2:-"staged staged-serial fused"}
if [ -z "$SKIP_TESTS" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/test_cpra_nccl.py > gpurun_out/r02_cpra_nccl_$N.log 2>&1
echo "nccl parity rc=$?"; grep -c OK gpurun_out/r02_cpra_nccl_$N.log; grep -v " OK " gpurun_out/r02_cpra_nccl_$N.log | tail -5
fi
for m in $MODES; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --exchange $m --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ab_${m}_$N.json 2> gpurun_out/r02_ab_${m}_$N.err
echo "bench $m rc=$?"; tail -3 gpurun_out/r02_ab_${m}_$N.err
python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r02_ab_${m}_$N.json') if x.startswith('{')][-1])
    print('$m', 'ms', round(l['ms_per_step'],3), 'G/s', round(l['value']/1e9,1), l.get('cpra_ms_per_step'), l.get('kernel_ms_per_step'))
except Exception as e: print('$m', e)
PY
done
