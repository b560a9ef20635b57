#!/bin/bash
# peer scatter: TMA bulk stores (HJB_PEER_BULK=1, default) against per-lane remote stores (=0); parity first
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
for b in 1 0; do
  HJB_PEER_BULK=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/test_cpra_nccl.py > gpurun_out/cpra_nccl_${N}_b$b.log 2>&1
  echo "bulk=$b parity:"; grep -E "CPRA_NCCL|Error|error|mismatch" gpurun_out/cpra_nccl_${N}_b$b.log | head -5
  HJB_PEER_BULK=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${N}_b$b.log 2>&1
  python - <<PY
import json
for ln in open("gpurun_out/bench_${N}_b$b.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("bulk=$b", round(d["ms_per_step"], 3), "ms", d["cpra_ms_per_step"], d["nvlink"]["achieved_gbs_per_direction"])
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_1.log 2>&1
python - <<PY
import json
for ln in open("gpurun_out/bench_1.log"):
    if ln.startswith("{"):
        d = json.loads(ln); print("N=1", round(d["ms_per_step"], 3), "ms", d.get("kernels_ms_per_step"))
PY
