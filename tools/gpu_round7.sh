#!/bin/bash
# 2-GPU call: sharded data-parallel parity + scaling bench
mkdir -p gpurun_out
nvidia-smi -L
echo "=== ddp parity fp32 wire ==="
COMM_DTYPE=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_parity.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -12
echo "=== ddp parity bf16 wire ==="
COMM_DTYPE=bf16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/ddp_parity.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -8
echo "=== bench 2 GPUs ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -3 | tee gpurun_out/bench_2gpu.json | cut -c1-600
echo "=== split-K wgrad probe (1 GPU) ==="
python tools/gpu_probe.py --only gemm_bench_wgrad_splitk --timeout 200 2>&1 | tail -2
echo "=== bench 1 GPU (mbs 4, split-K wgrad) ==="
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r7.json | cut -c1-300
