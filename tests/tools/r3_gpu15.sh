#!/bin/bash
# round 2, session 3: compute-sanitizer racecheck + synccheck over the new kernels on small inputs
cd "$(dirname "$0")/../.."
O=gpurun_out/r3n
mkdir -p $O
( timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tests/tools/sanitize_small.py 2>&1 | tail -40 ) > $O/racecheck.log
( timeout 400 compute-sanitizer --tool synccheck --print-limit 20 python tests/tools/sanitize_small.py 2>&1 | tail -30 ) > $O/synccheck.log
