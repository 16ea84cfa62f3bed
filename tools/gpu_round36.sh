#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels tests ==="
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
echo "=== attention bwd v3 vs v5 ==="
timeout 300 python tools/gpu_probe.py --only attn_bwd_experiments > /dev/null 2>&1
grep attn_bwd_experiments gpurun_out/probe.jsonl | tail -1
