#!/bin/bash
# round 2, session 3, GPU call 6: dominated-run shortcut only where a scan would run (A/B against the build without it),
# skewed Myers backtrace with an asynchronous next-window prefetch
cd "$(dirname "$0")/../.."
O=gpurun_out/r3f
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q -x 2>&1 | tail -6 ) > $O/pytest_aligner.log
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x -k "c3 or C3 or adaptive or golden or end2end or msa" 2>&1 | tail -6 ) > $O/pytest_poa.log
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
V=$PWD/build/variants/libgwb200_nopure.so
run c3_pure x=1 -- c3 1600 --factor 3 --allmem
run c3_nopure GWB200_LIB_PATH=$V -- c3 1600 --factor 3 --allmem
run c2_pure x=1 -- c2 1024
run c2_nopure GWB200_LIB_PATH=$V -- c2 1024
( GWB200_MYERS_TIMERS=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_skew.err | tail -1 ) > $O/c4_skew.json
