#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2h
mkdir -p $O gpurun_out/ncu
( timeout 900 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_spoa_accurate.py tests/test_gpu_global_aligners.py -m gpu -q 2>&1 | tail -12 ) > $O/pytest.log
q() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
q c3_bulk x=1 -- c3 1623 --factor 3 --allmem
q c3_nobulk GWB200_POA_BULK=0 -- c3 1623 --factor 3 --allmem
q c2_bulk x=1 -- c2 1024
q c2_nobulk GWB200_POA_BULK=0 -- c2 1024
q c2_nobulk_4096 GWB200_POA_BULK=0 -- c2 4096
( timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | tail -1 ) > $O/c4.json
( timeout 900 python bench.py --steps 3 --warmup 3 2> $O/bench.err | tail -1 ) > $O/bench.json
( timeout 900 python bench.py --impl reference --steps 4 --warmup 1 --no-extras 2> $O/bench_ref.err | tail -1 ) > $O/bench_ref.json
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none"
timeout 300 ncu $M -k regex:myers_banded -c 1 -f -o gpurun_out/ncu/r2h_traffic_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_c4.log 2>&1
