#!/bin/bash
# coverage evidence: MoE (C4 shape) rates + ncu of the grouped expert GEMM; ncu of the HBM-bound kernels at C2 sizes
mkdir -p gpurun_out
echo "=== MoE C4 ==="
timeout 600 python tools/gpu_probe.py --only moe_bench_c4 > /dev/null 2>&1
grep '"case": "moe_bench_c4"' gpurun_out/probe.jsonl | tail -1 | cut -c1-600
timeout 300 python tools/gpu_probe.py --only moe_layer_fwd_c4 > /dev/null 2>&1
grep '"case": "moe_layer_fwd_c4"' gpurun_out/probe.jsonl | tail -1 | cut -c1-300
echo "=== ncu grouped expert GEMM (c_fc forward) ==="
timeout 400 ncu --set full --clock-control none -k regex:gemm_bf16_kernel -s 4 -c 1 -f -o gpurun_out/prof_gemm_grouped_moe python tools/gpu_probe.py --case moe_layer_fwd_c4 > /dev/null 2>&1; echo rc=$?
echo "=== ncu HBM kernels ==="
PROBE_ITERS=1 PROBE_WARMUP=1 timeout 600 ncu --set full --clock-control none -k 'regex:rmsnorm_fwd_kernel|rmsnorm_bwd_kernel|rope_kernel|swiglu_fwd_kernel|swiglu_bwd_kernel|colsum_kernel|adamw_kernel|sumsq_kernel' -c 20 -f -o gpurun_out/prof_hbm_kernels python tools/gpu_probe.py --case elementwise_bench_c2 > /dev/null 2>&1; echo rc=$?
PROBE_ITERS=1 PROBE_WARMUP=0 timeout 400 ncu --set full --clock-control none -k regex:ce_fwd_bwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_ce python tools/gpu_probe.py --case embedding_ce > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
