#!/bin/bash
# round 2, session 3: Myers records only for blocks inside the band (parity, timing, traffic); ncu source counters of the C2 kernel
cd "$(dirname "$0")/../.."
O=gpurun_out/r3j
mkdir -p $O gpurun_out/ncu
( timeout 900 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q -x 2>&1 | tail -8 ) > $O/pytest_aligner.log
( GWB200_MYERS_TIMERS=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4.err | tail -1 ) > $O/c4.json
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none"
timeout 300 ncu $M -k regex:myers_banded -c 1 -f -o gpurun_out/ncu/r3j_traffic_c4 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_t_c4.log 2>&1
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --import-source on --clock-control none -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r3j_c2 \
    python tests/tools/quick_poa.py c2 1024 > $O/ncu_c2.log 2>&1
ncu -i gpurun_out/ncu/r3j_c2.ncu-rep --page source --csv > gpurun_out/ncu/r3j_c2_source.csv 2>/dev/null
