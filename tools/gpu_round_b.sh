#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py -x -q -k "two_gpu" 2>&1 | tail -8
echo "=== bench --mode train (1 GPU)"
python bench.py --mode train --steps 3 --warmup 2 2>&1 | tail -3 | tee gpurun_out/bench_train_1gpu.json
echo "=== bench default, 2 GPUs"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_2gpu.json
