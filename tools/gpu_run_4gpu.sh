#!/bin/bash
# 4-GPU call at HEAD (defaults: cluster-launch-control GEMM grids, chunked attention CTA order):
#   NCCL parity incl. HSDP 2 x 2 (replicate x shard), C2 stage 3 at 4 GPUs and at 1 GPU of the same box (weak-scaling efficiency).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nccl_parity.py -m gpu -q > gpurun_out/k_nccl_parity.log 2>&1
echo "nccl parity rc=$?"; tail -3 gpurun_out/k_nccl_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/k_c2_4gpu.json 2> gpurun_out/k_c2_4gpu.err
echo "c2 4 gpus rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/k_c2_1gpu.json 2> gpurun_out/k_c2_1gpu.err
echo "c2 1 gpu rc=$?"
python - <<'PY'
import json
r = {}
for f in ("k_c2_4gpu", "k_c2_1gpu"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        r[f] = d
        print(f, round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms", "per-gpu", round(d["tokens_per_sec_per_gpu"]), "gemm", round(d["roofline"]["achieved"]), d["clocks"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
if len(r) == 2:
    print("efficiency at 4 GPUs:", r["k_c2_4gpu"]["tokens_per_sec_per_gpu"] / r["k_c2_1gpu"]["tokens_per_sec_per_gpu"])
PY
