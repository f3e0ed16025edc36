#!/bin/bash
# round 2, session 3, GPU call 5: dominated-run shortcut + far-predecessor wait (POA), leaner skewed Myers pass and stage
cd "$(dirname "$0")/../.."
O=gpurun_out/r3e
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x -k "c3 or C3 or adaptive or golden or end2end or msa" 2>&1 | tail -6 ) > $O/pytest_poa.log
( timeout 600 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q -x 2>&1 | tail -6 ) > $O/pytest_aligner.log
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
run c3_default x=1 -- c3 1600 --factor 3 --allmem
run c3_fence GWB200_POA_ROW_FENCE=1 -- c3 1600 --factor 3 --allmem
run c3_nobulk GWB200_POA_BULK=0 -- c3 1600 --factor 3 --allmem
run c2_default x=1 -- c2 1024
run c2_nobulk GWB200_POA_BULK=0 -- c2 1024
( GWB200_MYERS_TIMERS=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_skew.err | tail -1 ) > $O/c4_skew.json
( GWB200_MYERS_TIMERS=1 GWB200_MYERS_SKEW=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_classic.err | tail -1 ) > $O/c4_classic.json
