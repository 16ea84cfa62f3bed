#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== fwd v2 correctness + bench ==="
timeout 300 python tools/gpu_probe.py --only attn_fwd_v2 > /dev/null 2>&1
grep '"case": "attn_fwd_v2_correctness"' gpurun_out/probe.jsonl | tail -1 | cut -c1-200
timeout 300 python tools/gpu_probe.py --only attn_bench_c2_fwd_v2 > /dev/null 2>&1
grep '"case": "attn_bench_c2_fwd_v2"' gpurun_out/probe.jsonl | tail -1 | cut -c150-330
timeout 300 python tools/gpu_probe.py --case attn_bench_c2 2>&1 | tail -1 | cut -c150-330
echo "=== bench mbs 4 / 6 / 7 ==="
for m in 4 6 7; do timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mbs $m 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_mbs$m.json | cut -c1-160; done
