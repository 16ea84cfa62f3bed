#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_fwd_b python tools/gpu_probe.py --case attn_bench_c2 > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/prof_attn_fwd_b.ncu-rep
