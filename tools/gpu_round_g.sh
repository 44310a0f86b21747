#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_default_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_default_8gpu.json')); print('default', d['value'], d['ms_per_step'], 'train', d['train']['value'], d['train']['ms_per_step'], d['train']['collective'])"
$TR --master-port 29522 bench.py --gpus 8 --mode stress --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_stress_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_stress_8gpu.json')); print('stress', d['value'], d['ms_per_step'], d['decoder'], d['fork_gemms'])"
$TR --master-port 29523 bench.py --gpus 8 --mode search --steps 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_search_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_search_8gpu.json')); print('search', d['value'], d['ms_per_step'])"
python -m pytest tests/test_gpu_train.py tests/test_gpu_edges.py -q -k "two_gpu or second_device" 2>&1 | tail -3
