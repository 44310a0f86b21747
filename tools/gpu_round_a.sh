#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== ncu smoke (driver style)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== ncu full dec_scan at metric config"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_scan -c 1 -o gpurun_out/r2a_dec_scan -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -5
ls -la gpurun_out/
