#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 85 (1 GPU): training-mode dropout (elementwise kernels, attention-probability dropout inside the attention kernels,
# engine integration) against the oracle with the same masks; gemm_dynamic with the tile id fetched one tile ahead (probe +
# C2 bench A/B); the attention tests again (the kernels gained a dropout branch in their per-element paths).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropout.py -m gpu -q -x > gpurun_out/c85_dropout_tests.log 2>&1
echo "dropout tests rc=$?"; tail -n 25 gpurun_out/c85_dropout_tests.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or empty" > gpurun_out/c85_attn_tests.log 2>&1
echo "attn tests rc=$?"; tail -n 3 gpurun_out/c85_attn_tests.log
timeout 600 python tools/gpu_probe.py --only gemm_dynamic --out gpurun_out/c85_probe.jsonl > gpurun_out/c85_probe.log 2>&1
echo "probe rc=$?"; cut -c1-2600 gpurun_out/c85_probe.jsonl
for dyn in 1 0 1 0; do
  DOLO_OPTIONS=gemm_dynamic=$dyn timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/c85_bench_dyn${dyn}_$RANDOM.json 2> gpurun_out/c85_bench.err
  echo "bench dyn=$dyn rc=$?"; ls -t gpurun_out/c85_bench_dyn${dyn}_*.json | head -1 | xargs cut -c1-220
done
