#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests ==="
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -8
