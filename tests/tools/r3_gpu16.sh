#!/bin/bash
# round 2, session 3: indel steps of the Myers backtrace from lane 0's speculative evaluation (parity + timing)
cd "$(dirname "$0")/../.."
O=gpurun_out/r3p
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q -x 2>&1 | tail -8 ) > $O/pytest_aligner.log
( GWB200_MYERS_TIMERS=1 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4.err | tail -1 ) > $O/c4.json
( GWB200_MYERS_TIMERS=1 GWB200_MYERS_SKEW=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_classic.err | tail -1 ) > $O/c4_classic.json
