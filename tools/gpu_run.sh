#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call I (1 GPU): lean softmax warps of the pipelined attention backward (no division, warp-private statistics, packed
# FFMA2 / FADD2 / FMUL2, scale folded into the dK / dQ epilogues): parity, interleaved A/B/C timing, ncu of the new kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention or empty" > gpurun_out/i_attn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/i_attn_tests.log
rm -f gpurun_out/i_probe.jsonl
for c in attn_bwd_variants_hd80 attn_bwd_variants_hd64; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/i_probe.jsonl > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_kernel_v3' --launch-skip 1 -c 1 -f -o gpurun_out/r02_attn_bwd_lean_hd80 python tools/ncu_targets.py > gpurun_out/i_ncu.log 2>&1
tail -c 1200 gpurun_out/i_attn_tests.log
cat gpurun_out/i_probe.jsonl | cut -c1-1800
tail -3 gpurun_out/i_ncu.log
