#!/bin/bash
cd /root/repo
run() { timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], round(d['roofline']['decoder_step_us'],2), d['kernel_ms_per_step']['gemm'], d['kernel_ms_per_step']['dec_scan'], round(d['e2e']['value']))"; }
LVSR_NO_F16_GEMM=1 run "nof16"
run "f16 lazy"
LVSR_NO_F16_GEMM=1 run "nof16"
run "f16 lazy"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_metric.py tests/test_gpu_search.py tests/test_gpu_train.py -q 2>&1 | tail -2
