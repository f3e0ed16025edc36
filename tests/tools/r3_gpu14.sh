#!/bin/bash
# round 2, session 3: compute-sanitizer memcheck over the new kernels on small inputs
cd "$(dirname "$0")/../.."
O=gpurun_out/r3n
mkdir -p $O
( timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tests/tools/sanitize_small.py 2>&1 | tail -40 ) > $O/memcheck.log
