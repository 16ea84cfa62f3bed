#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 87 (1 GPU): weight gradients on a second stream (DOLO_OVERLAP_WGRADS=1): model / full-width parity with the overlap on,
# C2 bench A/B; attention CTA order A/B and the overlap probe if the 2-GPU call did not run them.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or empty" > gpurun_out/c87_attn_tests.log 2>&1
echo "attention tests (chunked CTA order) rc=$?"; tail -n 2 gpurun_out/c87_attn_tests.log
DOLO_OVERLAP_WGRADS=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropout.py -m gpu -q -x > gpurun_out/c87_overlap_tests.log 2>&1
echo "model tests with overlapped wgrads rc=$?"; tail -n 4 gpurun_out/c87_overlap_tests.log | cut -c1-400
rm -f gpurun_out/c87_probe.jsonl
for c in attn_order_c2 attn_order_c5 attn_order_c4; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/c87_probe.jsonl > gpurun_out/c87_probe.log 2>&1
done
cut -c1-1600 gpurun_out/c87_probe.jsonl
for ov in 1 0 1 0; do
  DOLO_OVERLAP_WGRADS=$ov timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/c87_bench_ov${ov}_$RANDOM.json 2> gpurun_out/c87_bench.err
  echo "bench overlap=$ov rc=$?"; ls -t gpurun_out/c87_bench_ov${ov}_*.json | head -1 | xargs cut -c1-200
done
