#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2f
mkdir -p $O
q() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
q c3_full x=1 -- c3 1480 --factor 3 --allmem
q c3_full_1628 x=1 -- c3 1628 --factor 3 --allmem
q c3_k1 x=1 -- c3 148 --factor 3
q c2_1024 x=1 -- c2 1024
q c2_4096 x=1 -- c2 4096
q c2_v2_1024 GWB200_POA_KERNEL=v2 -- c2 1024
( timeout 1700 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -40 ) > $O/pytest_all.log
