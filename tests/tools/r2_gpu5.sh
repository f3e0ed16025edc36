#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2e
mkdir -p $O gpurun_out/ncu
q() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -7 ) >> $O/ab.txt; }
for g in 1 2 4; do
q c3_g${g}_full GWB200_POA_GROUP=$g -- c3 1480 --factor 3 --allmem
q c3_g${g}_k1 GWB200_POA_GROUP=$g -- c3 148 --factor 3
q c2_g${g}_1024 GWB200_POA_GROUP=$g -- c2 1024
q c2_g${g}_4096 GWB200_POA_GROUP=$g -- c2 4096
done
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --import-source on --clock-control none -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r2e_v3g2_4k \
    python tests/tools/quick_poa.py c3 1480 --len 4000 --factor 3 > gpurun_out/ncu/r2e_v3g2_4k_app.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_global_aligners.py -m gpu -q -x 2>&1 | tail -15 ) > $O/pytest_global.log
