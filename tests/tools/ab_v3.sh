#!/bin/bash
# Development aid (GPU box): A/B of the POA kernel generations and of the v3 switches, logs under gpurun_out/.
# usage: ab_v3.sh [quick|full]
cd "$(dirname "$0")/../.."
O=gpurun_out/ab_v3
mkdir -p $O
MODE=${1:-quick}
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" | tee -a $O/summary.txt
  ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -12 ) | tee -a $O/summary.txt
}
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv | tee $O/summary.txt
run c3_v3_ref      GWB200_POA_CTAS_PER_SM=10 -- c3 148 --factor 3 --ref
run c2_v3_ref      x=1 -- c2 1024 --ref
run c3_v3_full     x=1 -- c3 1480 --factor 3 --allmem
run c2_v3_4096     x=1 -- c2 4096
if [ "$MODE" = "full" ]; then
  run c3_v3_nobulk   GWB200_POA_CTAS_PER_SM=10 GWB200_POA_BULK=0 -- c3 148 --factor 3
  run c3_v3_notma    GWB200_POA_CTAS_PER_SM=10 GWB200_POA_TB_TMA=0 -- c3 148 --factor 3
  run c3_v2          GWB200_POA_KERNEL=v2 -- c3 148 --factor 3
  run c2_v2          GWB200_POA_KERNEL=v2 -- c2 1024
  run c3_v3_full_nb  GWB200_POA_BULK=0 -- c3 1480 --factor 3 --allmem
  run c3_v3_full_nt  GWB200_POA_TB_TMA=0 -- c3 1480 --factor 3 --allmem
  run c3_v2_full     GWB200_POA_KERNEL=v2 -- c3 1036 --factor 3 --allmem
fi
