#!/bin/bash
# round 2, session 3: the whole GPU suite once more on the final HEAD
cd "$(dirname "$0")/../.."
O=gpurun_out/r3q
mkdir -p $O
( timeout 330 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > $O/pytest_all.log
