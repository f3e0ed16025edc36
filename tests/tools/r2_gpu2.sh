#!/bin/bash
# Development aid (GPU box), round 2 call 2: wavefront rows (v4) correctness + A/B.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2b
mkdir -p $O
q() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -9 ) >> $O/ab.txt; }
( timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tests/tools/quick_poa.py c3 4 --len 2000 --factor 3 2>&1 | tail -40 ) > $O/sanitizer.txt
q c3_wf_small x=1 -- c3 16 --len 2000 --factor 3 --ref
q c3_v3_small GWB200_POA_WAVEFRONT=0 -- c3 16 --len 2000 --factor 3
q c3_wf_ref x=1 -- c3 148 --factor 3 --ref
q c3_v3_148 GWB200_POA_WAVEFRONT=0 -- c3 148 --factor 3
q c3_wf_full x=1 -- c3 888 --factor 3 --allmem
q c3_wf_full7 GWB200_POA_CTAS_PER_SM=7 -- c3 1036 --factor 3 --allmem
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x 2>&1 | tail -15 ) > $O/pytest_poa.log
