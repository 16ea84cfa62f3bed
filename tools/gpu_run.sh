#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 89 (1 GPU): split-softmax forward with FOUR threads per query row (attn_fwd_split = 3) against two threads per row and
# the single-buffer kernel: attention + dropout tests (every forward kernel), interleaved timing at the C2 / C5 / C4 shapes.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropout.py -m gpu -q -x -k "attention or empty" > gpurun_out/c89_attn_tests.log 2>&1
echo "attention tests rc=$?"; tail -n 3 gpurun_out/c89_attn_tests.log | cut -c1-300
rm -f gpurun_out/c89_probe.jsonl
for c in attn_fwd_kernels_c2 attn_fwd_kernels_c5 attn_fwd_kernels_c4; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/c89_probe.jsonl > gpurun_out/c89_probe.log 2>&1
done
cut -c1-1200 gpurun_out/c89_probe.jsonl; tail -c 400 gpurun_out/c89_probe.log
