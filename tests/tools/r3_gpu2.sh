#!/bin/bash
# round 2, session 3, GPU call 2: aligner parity with the skewed Myers score pass + C4 A/B against the classic passes
cd "$(dirname "$0")/../.."
O=gpurun_out/r3b
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_aligner.py tests/test_gpu_cpp_api.py -m gpu -q -x 2>&1 | tail -15 ) > $O/pytest_aligner.log
( timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 ) > $O/c4_skew.json
( GWB200_MYERS_SKEW=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 ) > $O/c4_classic.json
( GWB200_MYERS_SPECULATE=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 ) > $O/c4_skew_nospec.json
