#!/bin/bash
# round 2, session 3, GPU call 3: branch-free skewed Myers pass: parity, C4 A/B with phase timers, one ncu capture
cd "$(dirname "$0")/../.."
O=gpurun_out/r3c
mkdir -p $O gpurun_out/ncu
( timeout 900 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q -x 2>&1 | tail -15 ) > $O/pytest_aligner.log
( GWB200_MYERS_TIMERS=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_skew.err | tail -1 ) > $O/c4_skew.json
( GWB200_MYERS_TIMERS=1 GWB200_MYERS_SKEW=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_classic.err | tail -1 ) > $O/c4_classic.json
( GWB200_MYERS_TIMERS=1 GWB200_MYERS_SPECULATE=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_skew_nospec.err | tail -1 ) > $O/c4_skew_nospec.json
( GWB200_MYERS_TIMERS=1 GWB200_MYERS_SPECULATE=0 GWB200_MYERS_SKEW=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_classic_nospec.err | tail -1 ) > $O/c4_classic_nospec.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:myers_banded -c 1 -f -o gpurun_out/ncu/r3c_myers_skew \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_c4.log 2>&1
ncu -i gpurun_out/ncu/r3c_myers_skew.ncu-rep --page details --csv > gpurun_out/ncu/r3c_myers_skew_details.csv 2>/dev/null
ncu -i gpurun_out/ncu/r3c_myers_skew.ncu-rep --page source --csv > gpurun_out/ncu/r3c_myers_skew_source.csv 2>/dev/null
