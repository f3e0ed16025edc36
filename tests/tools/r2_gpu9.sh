#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2i
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv > $O/gpus.txt
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2> $O/c3_2gpu.err | tail -1 ) > $O/c3_2gpu.json
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2> $O/c4_2gpu.err | tail -1 ) > $O/c4_2gpu.json
( timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -4 ) > $O/pytest_sharded_nccl.log
