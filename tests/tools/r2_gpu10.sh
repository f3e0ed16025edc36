#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2j
mkdir -p $O
( timeout 400 python -m pytest tests/test_gpu_aligner.py "tests/test_gpu_poa.py::test_empty_group_stays_in_the_batch_without_touching_other_windows" -m gpu -q 2>&1 | tail -6 ) > $O/pytest.log
( timeout 900 python tests/tools/sweep_poa.py --ref --max-seconds 700 2> $O/sweep.err ) > $O/sweep_c5.jsonl
