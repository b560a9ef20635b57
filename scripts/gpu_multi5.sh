#!/bin/bash
# CPRA exchange modes at N GPUs: parity (nccl, fused, overlap), then bench fused vs overlap
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/test_cpra_nccl.py > gpurun_out/cpra_nccl_$N.log 2>&1
grep -E "CPRA_NCCL|Error|error|MISMATCH" gpurun_out/cpra_nccl_$N.log | head -8
for x in overlap fused; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --exchange $x > gpurun_out/bench_${N}_$x.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_${N}_$x.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("N=$N $x", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e9, 1), "Gtuples/s", d["cpra_ms_per_step"], d["nvlink"]["achieved_gbs_per_direction"])
PY
grep -E "Error|error|wrong" gpurun_out/bench_${N}_$x.log | head -3
done
