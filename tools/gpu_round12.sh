#!/bin/bash
# 2-GPU call: sharded runtime parity (wire dtype x accumulation x reshard) + 2-GPU bench with the comm CTA budget
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== ddp parity fp32 ==="
COMM_DTYPE=fp32 timeout 300 $TR --master-port 29541 tools/ddp_parity.py 2>&1 | grep -E "step|DDP_PARITY|Error|error" | tee gpurun_out/ddp_parity_fp32.log
echo "=== ddp parity bf16 ==="
COMM_DTYPE=bf16 timeout 300 $TR --master-port 29542 tools/ddp_parity.py 2>&1 | grep -E "step|DDP_PARITY|Error|error" | tee gpurun_out/ddp_parity_bf16.log
echo "=== ddp parity fp32 accum 2 ==="
COMM_DTYPE=fp32 ACCUM=2 timeout 300 $TR --master-port 29543 tools/ddp_parity.py 2>&1 | grep -E "step|DDP_PARITY|Error|error" | tee gpurun_out/ddp_parity_fp32_accum2.log
echo "=== ddp parity fp32 reshard ==="
COMM_DTYPE=fp32 RESHARD=1 timeout 300 $TR --master-port 29544 tools/ddp_parity.py 2>&1 | grep -E "step|DDP_PARITY|Error|error" | tee gpurun_out/ddp_parity_fp32_reshard.log
echo "=== bench 2 GPUs (comm ctas 8) ==="
timeout 900 $TR --master-port 29545 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_2gpu_ctas8.json
echo "=== bench 2 GPUs (comm ctas 16) ==="
DOLO_COMM_CTAS=16 timeout 900 $TR --master-port 29546 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_2gpu_ctas16.json
echo "=== bench 2 GPUs (comm ctas 4) ==="
DOLO_COMM_CTAS=4 timeout 900 $TR --master-port 29547 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_2gpu_ctas4.json
