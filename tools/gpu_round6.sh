#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu (all) ==="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench full (attn bwd v2 default) ==="
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r6.json | cut -c1-400
echo "=== bench mbs 4 ==="
timeout 900 python bench.py --steps 4 --warmup 3 --mbs 4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r6_mbs4.json | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv
echo "=== ncu full captures ==="
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_attn_fwd python tools/gpu_probe.py --case attn_bench_c2 > /dev/null 2>&1; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel_v2 -s 2 -c 1 -f -o gpurun_out/prof_attn_bwd_v2 python tools/gpu_probe.py --case attn_bench_c2 > /dev/null 2>&1; echo rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_fc python tools/gpu_probe.py --case gemm_bench_fc > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
