#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2k
mkdir -p $O gpurun_out/ncu
( timeout 600 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q 2>&1 | tail -8 ) > $O/pytest.log
( timeout 300 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | tail -1 ) > $O/c4_spec.json
( GWB200_MYERS_SPECULATE=0 timeout 300 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | tail -1 ) > $O/c4_nospec.json
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r2k_c2 \
    python tests/tools/quick_poa.py c2 1024 > $O/ncu_c2.log 2>&1
