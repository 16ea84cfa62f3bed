#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_2gpu.json | cut -c1-160
COMM_DTYPE=bf16 ACCUM=2 RESHARD=1 timeout 300 $TR --master-port 29572 tools/ddp_parity.py 2>&1 | grep -E "step|DDP_PARITY" | tee gpurun_out/ddp_parity_final.log
