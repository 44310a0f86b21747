#!/bin/bash
cd /root/repo
run() { timeout 60 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['ms_per_step'], round(d['roofline']['decoder_step_us'],2))"; }
LVSR_WS_SHIFT_KB=4096 run "tf32 shift 4MB"
LVSR_WS_SHIFT_KB=6144 run "tf32 shift 6MB"
LVSR_WS_SHIFT_KB=10240 run "tf32 shift 10MB"
LVSR_F16_GEMM=1 timeout 60 python bench.py --mode stress --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stress f16', d['value'], d['ms_per_step'], d['decoder']['us_per_step'], d['kernel_ms_per_step']['gemm'])"
