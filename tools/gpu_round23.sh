#!/bin/bash
# 8-GPU call: sharded parity at world 8 + the scaling point the driver will also run
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "=== ddp parity bf16 wire, world 8 ==="
COMM_DTYPE=bf16 timeout 300 $TR --master-port 29551 tools/ddp_parity.py 2>&1 | grep -E "step|DDP_PARITY|Error|error" | tee gpurun_out/ddp_parity_bf16_w8.log
echo "=== bench 8 GPUs ==="
timeout 600 $TR --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_8gpu.json | cut -c1-400
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
