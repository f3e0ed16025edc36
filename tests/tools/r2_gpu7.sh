#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2g
mkdir -p $O gpurun_out/ncu
( timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_global_aligners.py tests/test_gpu_spoa_accurate.py tests/test_gpu_cpp_api.py tests/test_pyshim.py -m gpu -q --durations=5 2>&1 | tail -40 ) > $O/pytest_new.log
( timeout 900 python bench.py --steps 5 --warmup 3 2> $O/bench.err | tail -2 ) > $O/bench.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $O/bench_ref.err | tail -2 ) > $O/bench_ref.json
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none"
timeout 600 ncu $M -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r2g_traffic_c3 python tests/tools/quick_poa.py c3 592 --factor 3 > $O/ncu_c3.log 2>&1
timeout 300 ncu $M -k regex:poa_window_kernel_v3 -c 1 -f -o gpurun_out/ncu/r2g_traffic_c2 python tests/tools/quick_poa.py c2 1024 > $O/ncu_c2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:myers_banded -c 1 -f -o gpurun_out/ncu/r2g_myers_c4 \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_c4.log 2>&1
