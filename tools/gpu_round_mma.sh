#!/bin/bash
# BiGRU tensor-core kernel: parity, A/B against the FFMA kernel, section trace
cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_metric.py tests/test_gpu_train.py tests/test_gpu_edges.py -x -q 2>&1 | tail -5
for m in 0 1; do
  echo "== LVSR_BIGRU_MMA=$m"
  LVSR_BIGRU_MMA=$m python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d.get('kernel_ms_per_step')))"
done
LVSR_BIGRU_TRACE=1 python bench.py --steps 1 --warmup 3 2>&1 | grep "bigru trace" | tail -4
