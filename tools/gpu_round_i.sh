#!/bin/bash
# round 2, snapshot i: tensor-core BiGRU -- new edge tests, launch list, ncu --set full of the kernel
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py -x -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2i_metric.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/r2i_launch_bench.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bigru_mma -s 1 -c 1 -f -o gpurun_out/r2i_bigru_mma \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/r2i_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -5
