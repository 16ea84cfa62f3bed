#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== elementwise bandwidth ==="
timeout 300 python tools/gpu_probe.py --only elementwise_bench_c2 > /dev/null 2>&1
grep '"case": "elementwise_bench_c2"' gpurun_out/probe.jsonl | tail -1 | cut -c1-700
echo "=== attention ==="
timeout 300 python tools/gpu_probe.py --case attn_bench_c2 2>&1 | tail -1 | cut -c150-330
timeout 300 python tools/gpu_probe.py --case attn_bench_hd128 2>&1 | tail -1 | cut -c1-900
echo "=== bench ==="
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-step gpurun_out/step_profile_mbs6.json 2>&1 | tail -1 | tee gpurun_out/bench_1gpu.json | cut -c1-200
