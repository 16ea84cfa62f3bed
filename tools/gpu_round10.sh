#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu ==="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== ddp parity (2 ranks) ==="
COMM_DTYPE=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_parity.py > gpurun_out/ddp_fp32.log 2>&1; grep -E "^step|DDP_PARITY|Error" gpurun_out/ddp_fp32.log | head -8
COMM_DTYPE=bf16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/ddp_parity.py > gpurun_out/ddp_bf16.log 2>&1; grep -E "^step|DDP_PARITY|Error" gpurun_out/ddp_bf16.log | head -8
echo "=== bench 1 GPU (CTA-pair GEMM default) ==="
timeout 900 python bench.py --steps 6 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r10.json | cut -c1-300
echo "=== bench 2 GPUs ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep "^{" | tail -1 | tee gpurun_out/bench_2gpu.json | cut -c1-300
echo "=== reference arm on this box ==="
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-700
