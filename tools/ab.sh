#!/bin/bash
# A/B of two builds of the library on one box: tools/ab/liblvsr_old.so (copied before a change) vs the in-tree build
for v in old new old new; do
  if [ $v = old ]; then export LVSR_B200_LIB=$PWD/tools/ab/liblvsr_old.so; else unset LVSR_B200_LIB; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['roofline']['decoder_step_us'],2), d['kernel_ms_per_step']['bigru'], d['kernel_ms_per_step']['gemm'])"
done
