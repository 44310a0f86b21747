#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -5
python bench.py --mode train --steps 3 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_train_1gpu_c.json
python -c "
import json; d=json.load(open('gpurun_out/bench_train_1gpu_c.json')); print(d['ms_per_step'], json.dumps(d['kernel_ms_per_step']))"
echo "=== ncu gemm_tc (metric config: launches 2..4 = layer 1, 2, 3 fork GEMMs)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 2 -o gpurun_out/r2d_gemm_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train 2>&1 | grep -i "prof\|error" | tail -4
echo "=== ncu training kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"att_bwd_kernel|bigru_bwd_kernel" -s 2 -c 3 -o gpurun_out/r2d_train -f python bench.py --mode train --steps 1 --warmup 2 2>&1 | grep -i "prof\|error" | tail -4
ls -la gpurun_out/*.ncu-rep
