#!/bin/bash
# 4-GPU call: how many SMs does the sharded runtime have to leave to NCCL?  (persistent GEMM grids give up exactly that many)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for c in 2 4 8; do
  echo "=== bench 4 GPUs, comm ctas $c ==="
  DOLO_COMM_CTAS=$c timeout 600 $TR --master-port $((29560 + c)) bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_4gpu_ctas$c.json | cut -c1-150
done
