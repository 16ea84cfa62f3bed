#!/bin/bash
# Call 95 (2 GPUs, the round's last GPU minutes): sharded == unsharded over NCCL for MoE blocks at HEAD (lazily cleared expert
# gradients in the shared stage-3 gradient buffers; then the resident mode if time is left).
set -u
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      tools/ddp_parity.py > gpurun_out/c95_nccl_parity_$name.log 2>&1
  echo "$name rc=$?"; grep -E "^step|DDP_PARITY" gpurun_out/c95_nccl_parity_$name.log | cut -c1-200
}
run moe_stage3_fp32_accum2 MOE=1 COMM_DTYPE=fp32 RESHARD=1 ACCUM=2
run moe_resident_bf16_accum2 MOE=1 COMM_DTYPE=bf16 RESHARD=0 ACCUM=2
