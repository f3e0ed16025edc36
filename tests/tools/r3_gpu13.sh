#!/bin/bash
# round 2, session 3: two GPUs: NCCL sharded parity test, weak-scaling bench lines (C3, C4)
cd "$(dirname "$0")/../.."
O=gpurun_out/r3m
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv > $O/gpus.txt
( timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -4 ) > $O/pytest_sharded_nccl.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c3_2gpu.err | tail -1 ) > $O/c3_2gpu.json
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_2gpu.err | tail -1 ) > $O/c4_2gpu.json
