#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 91 (1 GPU): RoPE with native packed bf16 arithmetic -- bit-exactness against the oracle (kernel + model tests) and GB/s.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "rope or logits or golden or yarn" > gpurun_out/c91_rope_tests.log 2>&1
echo "rope / model tests rc=$?"; tail -n 4 gpurun_out/c91_rope_tests.log | cut -c1-300
rm -f gpurun_out/c91_probe.jsonl
timeout 300 python tools/gpu_probe.py --only rope --out gpurun_out/c91_probe.jsonl > gpurun_out/c91_probe.log 2>&1
timeout 300 python tools/gpu_probe.py --only elementwise_bench_c2 --out gpurun_out/c91_probe.jsonl >> gpurun_out/c91_probe.log 2>&1
cut -c1-1500 gpurun_out/c91_probe.jsonl; tail -c 300 gpurun_out/c91_probe.log
