#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/ncu gpurun_out/r2c
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --import-source on --clock-control none -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r2c_v4_4k \
    python tests/tools/quick_poa.py c3 296 --len 4000 --factor 3 > gpurun_out/ncu/r2c_v4_4k_app.log 2>&1
( timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | tail -3 ) > gpurun_out/r2c/c4.txt
( timeout 600 python -m pytest tests/test_gpu_aligner.py -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/r2c/pytest_aligner.log
