#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== attention bwd experiments ==="
timeout 300 python tools/gpu_probe.py --only attn_bwd_experiments > /dev/null 2>&1
grep attn_bwd_experiments gpurun_out/probe.jsonl | tail -1
timeout 300 python tools/gpu_probe.py --only attn_bench_c2_fwd_v2 > /dev/null 2>&1
grep '"case": "attn_bench_c2_fwd_v2"' gpurun_out/probe.jsonl | tail -1 | cut -c1-900
timeout 300 python tools/gpu_probe.py --case attn_bench_c2 2>&1 | tail -1 | cut -c1-900
echo "=== gemm pair bench ==="
timeout 300 python tools/gpu_probe.py --only gemm_pair_bench > /dev/null 2>&1
grep '"case": "gemm_pair_bench"' gpurun_out/probe.jsonl | tail -1 | cut -c1-1500
echo "=== bench ==="
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-step gpurun_out/step_profile_mbs4.json 2>&1 | tail -1 | tee gpurun_out/bench_1gpu.json | cut -c1-400
