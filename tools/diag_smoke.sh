#!/bin/bash
# plain, then under ncu the way the driver lists launches
mkdir -p gpurun_out
NCU="ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/diag_launches.csv"
echo "=== plain"; python tools/diag_smoke.py 2>&1 | tail -20
echo "=== ncu"; $NCU python tools/diag_smoke.py 2>&1 | tail -20
echo "=== ncu stepwise decoder"; LVSR_NO_DEC_SCAN=1 $NCU python tools/diag_smoke.py 2>&1 | tail -20
echo "=== ncu no tc gemm"; LVSR_NO_TC_GEMM=1 $NCU python tools/diag_smoke.py 2>&1 | tail -20
echo "=== ncu smoke exactly"; $NCU python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
