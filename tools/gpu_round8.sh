#!/bin/bash
mkdir -p gpurun_out
echo "=== attention v3 + fwd diet ==="
python tools/gpu_probe.py --only attn_v3 --timeout 150 2>&1 | tail -2 | cut -c1-700
python tools/gpu_probe.py --only attn_bench_c2 --timeout 150 2>&1 | tail -4 | cut -c1-200
echo "=== pytest gpu ==="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== ddp parity (2 ranks) ==="
COMM_DTYPE=fp32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_parity.py > gpurun_out/ddp_fp32.log 2>&1; grep -E "step|DDP_PARITY|Error|error" gpurun_out/ddp_fp32.log | head -12
COMM_DTYPE=bf16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/ddp_parity.py > gpurun_out/ddp_bf16.log 2>&1; grep -E "step|DDP_PARITY|Error|error" gpurun_out/ddp_bf16.log | head -12
echo "=== bench 1 GPU ==="
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r8.json | cut -c1-300
echo "=== bench 1 GPU attn bwd v3 ==="
DOLO_ATTN_BWD=3 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r8_v3.json | cut -c1-300
echo "=== bench 2 GPUs ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep "^{" | tail -1 | tee gpurun_out/bench_2gpu.json | cut -c1-300
echo "=== ncu wgrad gemm ==="
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_wgrad python tools/gpu_probe.py --case gemm_bench_wgrad > /dev/null 2>&1; echo rc=$?
