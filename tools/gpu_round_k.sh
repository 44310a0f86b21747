#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_metric.py tests/test_gpu_search.py tests/test_gpu_edges.py tests/test_gpu_train.py -q 2>&1 | tail -3
for v in old new old new; do
  if [ $v = old ]; then export LVSR_B200_LIB=$PWD/tools/ab/liblvsr_old.so; else unset LVSR_B200_LIB; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), d['ms_per_step'], round(d['roofline']['decoder_step_us'],2), d['kernel_ms_per_step'])"
done
