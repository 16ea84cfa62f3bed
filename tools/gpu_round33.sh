#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== elementwise bandwidth ==="
timeout 300 python tools/gpu_probe.py --only elementwise_bench_c2 > /dev/null 2>&1
grep '"case": "elementwise_bench_c2"' gpurun_out/probe.jsonl | tail -1 | cut -c1-330
echo "=== bench (default flags) ==="
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-200
