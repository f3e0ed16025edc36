#!/bin/bash
# round 2, session 3, GPU call 1: POA parity on the 8-cells-per-lane int32 rows + A/B against the 4-cell build
cd "$(dirname "$0")/../.."
O=gpurun_out/r3a
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x 2>&1 | tail -8 ) > $O/pytest_poa.log
run() { # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt
  ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt
}
V=$PWD/build/variants/libgwb200_cpl4.so
run c3_cpl8_g2 x=1 -- c3 1600 --factor 3 --allmem
run c3_cpl8_g1 GWB200_POA_GROUP=1 -- c3 1600 --factor 3 --allmem
run c3_cpl4_g2 GWB200_LIB_PATH=$V -- c3 1600 --factor 3 --allmem
run c3_cpl8_k1 x=1 -- c3 148 --factor 3
run c2_new x=1 -- c2 1024
run c2_old GWB200_LIB_PATH=$V -- c2 1024
run c2_new_4096 x=1 -- c2 4096
run c2_old_4096 GWB200_LIB_PATH=$V -- c2 4096
