#!/bin/bash
# First GPU call of the next round: things built at the end of round 1 that could not be run on GPUs (no budget left).
#   gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_next_round_first.sh'
# 1. HSDP (2 shard x 2 replicate) parity against the unsharded run, fp32 and bf16 wire
# 2. communication CTA budget sweep at 4 GPUs (DOLO_COMM_CTAS = 2 / 4 / 8 / 16)
# 3. generate entry point smoke on one GPU
mkdir -p gpurun_out
N=${N:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
echo "=== HSDP parity (SHARD=2, world $N) ==="
SHARD=2 COMM_DTYPE=fp32 run 29541 tools/ddp_parity.py 2>&1 | tail -6 | tee gpurun_out/hsdp_parity_fp32.log
SHARD=2 COMM_DTYPE=bf16 ACCUM=2 run 29542 tools/ddp_parity.py 2>&1 | tail -6 | tee gpurun_out/hsdp_parity_bf16_accum2.log
echo "=== comm CTA sweep (world $N) ==="
for c in 2 4 8 16; do
  DOLO_COMM_CTAS=$c run $((29550 + c)) bench.py --gpus "$N" --steps 6 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${N}gpu_commctas${c}.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_${N}gpu_commctas${c}.json")); print("comm ctas $c:", d["value"], d["unit"], d["ms_per_step"], "ms")
PY
done
echo "=== generation tests + full GPU suite ==="
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
