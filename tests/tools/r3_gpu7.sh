#!/bin/bash
# round 2, session 3, GPU call 7 (final): the whole GPU suite, the bench lines, launch list and DRAM traffic of the shipped kernels
cd "$(dirname "$0")/../.."
O=gpurun_out/r3g
mkdir -p $O gpurun_out/ncu
( timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 ) > $O/pytest_all.log
( timeout 900 python bench.py --steps 3 --warmup 3 2> $O/bench.err | tail -1 ) > $O/bench_c3.json
( timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4.err | tail -1 ) > $O/bench_c4.json
( timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c2.err | tail -1 ) > $O/bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/launches_c3.log 2>&1
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none"
timeout 400 ncu $M -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r3g_traffic_c3 python tests/tools/quick_poa.py c3 592 --factor 3 > $O/ncu_t_c3.log 2>&1
timeout 300 ncu $M -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r3g_traffic_c2 python tests/tools/quick_poa.py c2 1024 > $O/ncu_t_c2.log 2>&1
timeout 300 ncu $M -k regex:myers_banded -c 1 -f -o gpurun_out/ncu/r3g_traffic_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_t_c4.log 2>&1
