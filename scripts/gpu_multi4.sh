#!/bin/bash
# CPRA at N GPUs: parity (fused + NCCL exchange, rows against the oracle), then the fused bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/test_cpra_nccl.py > gpurun_out/cpra_nccl_$N.log 2>&1
grep -E "CPRA_NCCL|Error|error|mismatch" gpurun_out/cpra_nccl_$N.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$N.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_$N.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("N=$N", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e9, 1), "Gtuples/s", d["cpra_ms_per_step"], d["nvlink"], "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e9, 2))
PY
tail -3 gpurun_out/bench_$N.log | cut -c1-300
