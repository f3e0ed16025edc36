#!/bin/bash
# round 2, session 3: the driver's own sequence on the shipped build: smoke(), then the reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out/r3l
mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -5 ) > $O/smoke.log
( timeout 1200 python bench.py --impl reference --steps 4 --warmup 1 2> $O/bench_ref.err | tail -1 ) > $O/bench_ref.json
