#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu (all) ==="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== smoke ==="
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== ncu launch list (2 layers) ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_L2.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches_L2.csv
echo "=== bench full ==="
timeout 900 python bench.py --steps 6 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_r4.json | cut -c1-1500
