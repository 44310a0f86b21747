#!/bin/bash
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/diag_launches.csv"
for i in 1 2 3; do echo "=== ncu run $i"; $NCU python tools/diag_smoke.py 2>&1 | grep -v "^==" | head -60; done
echo "=== plain"; python tools/diag_smoke.py 2>&1 | head -30
