#!/bin/bash
# peer-scatter grid experiment: is the fused scatter NVLink-bound (time independent of the number of CTAs)?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/test_cpra_nccl.py > gpurun_out/cpra_nccl_$N.log 2>&1; grep -E "CPRA_NCCL" gpurun_out/cpra_nccl_$N.log
for c in 0 96 64 48 32; do
  HJB_PEER_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${N}_c$c.log 2>&1
  python - <<PY
import json
for ln in open("gpurun_out/bench_${N}_c$c.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("peer_ctas=$c", round(d["ms_per_step"], 3), "ms", d["cpra_ms_per_step"], d["nvlink"]["achieved_gbs_per_direction"])
PY
done
