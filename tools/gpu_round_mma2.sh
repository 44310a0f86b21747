#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_metric.py -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d.get('kernel_ms_per_step')))"
LVSR_BIGRU_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 3 2>&1 | grep "bigru trace" | tail -8
