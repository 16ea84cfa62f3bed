#!/bin/bash
# Call 96 (2 GPUs, the round's last GPU seconds): the MoE cases + the dense stage-3 case of tests/test_nccl_parity.py at HEAD.
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_nccl_parity.py -q -m gpu -k "moe or stage3_reshard_bf16" -p no:cacheprovider > gpurun_out/c96_nccl_parity.log 2>&1
echo "rc=$?"; tail -n 5 gpurun_out/c96_nccl_parity.log | cut -c1-300
