#!/bin/bash
# 1-GPU call: full GPU test suite, HBM-kernel bandwidths at C2 sizes, in-situ per-kernel step profile, ncu captures of
# the attention kernels (source-level) and the fc GEMM (DRAM traffic for roofline.traffic)
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== elementwise bandwidth ==="
timeout 300 python tools/gpu_probe.py --only elementwise_bench_c2 > /dev/null 2>&1
grep elementwise_bench_c2 gpurun_out/probe.jsonl | tail -1
echo "=== bench (1 GPU, full) + step profile ==="
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-step gpurun_out/step_profile_mbs4.json 2>&1 | tail -2 | tee gpurun_out/bench_1gpu.json
echo "=== ncu attention bwd v3 ==="
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel_v3 -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd_v3 python tools/gpu_probe.py --case attn_bench_c2 > /dev/null 2>&1; echo rc=$?
echo "=== ncu attention fwd ==="
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_fwd python tools/gpu_probe.py --case attn_bench_c2 > /dev/null 2>&1; echo rc=$?
echo "=== ncu pair gemm fc ==="
timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_fc_pair python tools/gpu_probe.py --case gemm_bench_fc > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
