#!/bin/bash
# round 2, session 3: C5 sweep on the shipped build, every cell compared with the unmodified reference kernels
cd "$(dirname "$0")/../.."
O=gpurun_out/r3k
mkdir -p $O
( timeout 700 python tests/tools/sweep_poa.py --ref --max-seconds 540 2> $O/sweep.err ) > $O/sweep_c5.jsonl
