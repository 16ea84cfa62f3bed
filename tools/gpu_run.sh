#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call L (1 GPU): pipelined attention backward with uniform 16-column chunks (one N = head_dim MMA per K step for dV / dK / dQ)
# and the early release of the Q / dO stage: parity, interleaved A/B/C timing.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention or empty" > gpurun_out/l_attn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/l_attn_tests.log
rm -f gpurun_out/l_probe.jsonl
for c in attn_bwd_variants_hd80 attn_bwd_variants_hd64; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/l_probe.jsonl > /dev/null 2>&1
done
tail -c 1500 gpurun_out/l_attn_tests.log
cat gpurun_out/l_probe.jsonl | cut -c1-1800
