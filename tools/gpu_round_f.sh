#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== launch list of the default command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_metric.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/b_under_ncu.log 2>&1; tail -c 300 gpurun_out/b_under_ncu.log
echo "=== smoke under ncu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== default bench"
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_default_1gpu.json; python -c "
import json; d=json.load(open('gpurun_out/bench_default_1gpu.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['train']['ms_per_step'], d['cpu_baseline']['value'])"
