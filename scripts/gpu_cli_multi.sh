#!/bin/bash
# the reference's command line on N GPUs: ./write, then ./phj (one GPU) and HJB_GPUS=N ./cpra must agree
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
BIN=$GRAFT_REPO_ROOT/hash_join_codes_knl_b200/bin
W=$(mktemp -d)
cd $W
for sz in "4000000 1000000" "67108864 67108864"; do
  set -- $sz
  $BIN/write 4 $1 $2 > /dev/null || exit 1
  $BIN/phj 4 $1 $2 1 | tail -1 > phj.json
  HJB_GPUS=$N $BIN/cpra 4 $1 $2 > cpra.out 2> cpra.err; echo "cpra rc=$?"; head -2 cpra.out; tail -2 cpra.err
  tail -1 cpra.out > cpra.json
  HJB_GPUS=$N $BIN/cpra 4 $1 $2 | head -2 | tr '\n' ' '; echo "(second run)"
  python - <<'PY'
import json
a, b = json.load(open("phj.json")), json.load(open("cpra.json"))
keys = ("join_tuples", "sum_key", "sum_outer", "sum_inner")
print("outer", a["outer_tuples"], "inner", a["inner_tuples"], "phj == cpra:", all(a[k] == b[k] for k in keys), [b[k] for k in keys], "gpus", b["gpus"],
      "cpra seconds", b["seconds"], "phj seconds", a["seconds"])
PY
  rm -f *.txt
done
cd /; rm -rf $W
