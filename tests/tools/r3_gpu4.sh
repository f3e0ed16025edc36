#!/bin/bash
# round 2, session 3, GPU call 4: ring-only rows: parity + timing + ncu source counters of the C3 kernel at the bench shape
cd "$(dirname "$0")/../.."
O=gpurun_out/r3d
mkdir -p $O gpurun_out/ncu
( timeout 900 python -m pytest tests/test_gpu_poa.py -m gpu -q -x -k "c3 or C3 or adaptive or golden or end2end or msa" 2>&1 | tail -6 ) > $O/pytest_poa.log
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
run c3_ringonly x=1 -- c3 1600 --factor 3 --allmem
run c2_ringonly x=1 -- c2 1024
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --import-source on --clock-control none -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r3d_c3 \
    python tests/tools/quick_poa.py c3 1600 --factor 3 --allmem > $O/ncu_c3.log 2>&1
ncu -i gpurun_out/ncu/r3d_c3.ncu-rep --page source --csv > gpurun_out/ncu/r3d_c3_source.csv 2>/dev/null
ncu -i gpurun_out/ncu/r3d_c3.ncu-rep --page details --csv > gpurun_out/ncu/r3d_c3_details.csv 2>/dev/null
ls -la gpurun_out/ncu >> $O/ncu_c3.log
