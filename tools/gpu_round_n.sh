#!/bin/bash
# final validation of the snapshot: full GPU suite, default bench line, smoke
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py 2>gpurun_out/bench_r2k_1gpu.err | tail -1 > gpurun_out/bench_r2k_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2k_1gpu.json')); print('default', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['decoder_step_us'], d['roofline']['encoder_recurrence'], d['kernel_ms_per_step'], 'train', d['train']['value'], d['train']['ms_per_step'], d['clocks'], d['gpu_launches'])"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
