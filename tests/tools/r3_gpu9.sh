#!/bin/bash
# round 2, session 3: topological sort in rank space (A/B against the sort on node ids), POA parity
cd "$(dirname "$0")/../.."
O=gpurun_out/r3i
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x -k "c3 or C3 or adaptive or golden or end2end or msa" 2>&1 | tail -6 ) > $O/pytest_poa.log
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
run c3_ranked x=1 -- c3 1600 --factor 3 --allmem
run c3_by_id GWB200_POA_RANKED_SORT=0 -- c3 1600 --factor 3 --allmem
run c2_ranked x=1 -- c2 1024
run c2_by_id GWB200_POA_RANKED_SORT=0 -- c2 1024
