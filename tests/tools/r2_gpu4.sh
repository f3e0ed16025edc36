#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2d
mkdir -p $O
q() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
q c3_v3diet_full x=1 -- c3 1480 --factor 3 --allmem
q c3_v3diet_ref x=1 -- c3 148 --factor 3 --ref
q c3_v3diet_k8 GWB200_POA_CTAS_PER_SM=8 -- c3 1184 --factor 3 --allmem
q c3_v4c_148 GWB200_POA_WAVEFRONT=1 -- c3 148 --factor 3 --ref
q c3_v4c_full GWB200_POA_WAVEFRONT=1 -- c3 888 --factor 3 --allmem
q c3_v3diet_4k x=1 -- c3 1480 --len 4000 --factor 3
q c2_v3diet x=1 -- c2 1024 --ref
q c2_v3diet_4096 x=1 -- c2 4096
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x 2>&1 | tail -5 ) > $O/pytest_poa.log
