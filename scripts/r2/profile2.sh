#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python scripts/r2/sanity.py | tail -1
timeout 300 python scripts/r2/exp.py cfg1,cfg3 npj 1 2>&1 | tee gpurun_out/r02_exp_npj.log
cap() { local name=$1 k=$2 s=$3; shift 3
  HJB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c 1 -o gpurun_out/r02_prof_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1
}
cap npj_cfg1 'k_npj_probe' 2 python scripts/r2/exp.py cfg1 npj 1
cap npj_cfg3 'k_npj_probe' 2 python scripts/r2/exp.py cfg3 npj 1
