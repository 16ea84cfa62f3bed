#!/bin/bash
mkdir -p gpurun_out
echo "=== attention bwd v2 ==="
python tools/gpu_probe.py --only attn_v2 --timeout 120 2>&1 | tail -3 | cut -c1-600
python tools/gpu_probe.py --only attn_bench_c2 --timeout 150 2>&1 | tail -4 | cut -c1-200
echo "=== pytest gpu (all) ==="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
