#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels tests ==="
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
echo "=== attention bwd experiments (v3 vs v4) ==="
timeout 300 python tools/gpu_probe.py --only attn_bwd_experiments > /dev/null 2>&1
grep attn_bwd_experiments gpurun_out/probe.jsonl | tail -1
echo "=== bench v4 ==="
DOLO_ATTN_BWD=4 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v4.json | cut -c1-330
