#!/bin/bash
# Development aid (GPU box), round 2 call 1: GPU tests, v2/v3 A/B on C3, ncu captures of v3 (4 kb x 32, full residency) and Myers C4.
cd "$(dirname "$0")/../.."
O=gpurun_out/r2a
mkdir -p $O gpurun_out/ncu
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > $O/pytest.log
q() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "=== $name: ${envs[*]} $*" >> $O/ab.txt; ( env "${envs[@]}" timeout 300 python tests/tools/quick_poa.py "$@" 2>&1 | tail -8 ) >> $O/ab.txt; }
q c3_v2_full GWB200_POA_KERNEL=v2 -- c3 1036 --factor 3 --allmem
q c3_v3_full x=1 -- c3 1480 --factor 3 --allmem
q c3_v3_4k x=1 -- c3 1480 --len 4000 --factor 3
q c3_v2_4k GWB200_POA_KERNEL=v2 -- c3 1036 --len 4000 --factor 3
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SpeedOfLight --section InstructionStats --import-source on --clock-control none -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r2a_v3_4k \
    python tests/tools/quick_poa.py c3 1480 --len 4000 --factor 3 > gpurun_out/ncu/r2a_v3_4k_app.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:myers_banded -c 1 -f -o gpurun_out/ncu/r2a_myers_c4 \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu/r2a_myers_app.log 2>&1
ls -la gpurun_out/ncu >> $O/ab.txt
nvidia-smi --query-gpu=name,memory.used --format=csv >> $O/ab.txt
