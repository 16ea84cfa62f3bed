#!/bin/bash
mkdir -p gpurun_out
echo "=== kernels tests ==="
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
echo "=== fwd v2 correctness + bench ==="
timeout 300 python tools/gpu_probe.py --only attn_fwd_v2 > /dev/null 2>&1
grep '"case": "attn_fwd_v2_correctness"' gpurun_out/probe.jsonl | tail -1 | cut -c1-300
timeout 300 python tools/gpu_probe.py --only attn_bench_c2_fwd_v2 > /dev/null 2>&1
grep '"case": "attn_bench_c2_fwd_v2"' gpurun_out/probe.jsonl | tail -1 | cut -c1-400
echo "=== ncu attention bwd v3 (after fixes) ==="
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel_v3 -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd_v3b python tools/gpu_probe.py --case attn_bench_c2 > /dev/null 2>&1; echo rc=$?
