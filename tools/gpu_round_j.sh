#!/bin/bash
# round 2, snapshot j: full GPU suite + every bench mode on one B200 (files -> profiles/bench_r2j_*.json)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py 2>gpurun_out/bench_r2j_1gpu.err | tail -1 > gpurun_out/bench_r2j_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2j_1gpu.json')); print('default', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['kernel_ms_per_step'], 'train', d['train']['value'], d['train']['ms_per_step'])"
timeout 300 python bench.py --mode train --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r2j_train_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2j_train_1gpu.json')); print('train', d['value'], d['ms_per_step'], json.dumps(d['kernel_ms_per_step']))"
timeout 300 python bench.py --mode search --steps 2 2>/dev/null | tail -1 > gpurun_out/bench_r2j_search_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2j_search_1gpu.json')); print('search', d['value'], d['ms_per_step'])"
timeout 300 python bench.py --mode stress --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r2j_stress_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2j_stress_1gpu.json')); print('stress', d['value'], d['ms_per_step'])"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 > gpurun_out/bench_r2j_reference.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2j_reference.json')); print('reference', d.get('value'), d.get('unavailable'))"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
