#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 84 (1 GPU): cluster-launch-control tile scheduling of the GEMMs (gemm_dynamic): exactness, A/B timing, behaviour next
# to a kernel that holds 8 SMs; the GPU kernel / MoE / model tests once more with gemm_dynamic = 1; attention tests with the
# new default backward variant; C2 bench static vs dynamic.
set -u
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe.py --only gemm_dynamic --out gpurun_out/c84_probe.jsonl > gpurun_out/c84_probe.log 2>&1
echo "probe rc=$?"
cut -c1-3000 gpurun_out/c84_probe.jsonl
tail -c 600 gpurun_out/c84_probe.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or empty" > gpurun_out/c84_attn_tests.log 2>&1
echo "attn tests rc=$?"; tail -n 3 gpurun_out/c84_attn_tests.log
DOLO_OPTIONS=gemm_dynamic=1 timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_moe.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/c84_dynamic_tests.log 2>&1
echo "dynamic tests rc=$?"; tail -n 5 gpurun_out/c84_dynamic_tests.log
for dyn in 0 1 0 1; do
  DOLO_OPTIONS=gemm_dynamic=$dyn timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/c84_bench_dyn${dyn}.json 2> gpurun_out/c84_bench_dyn${dyn}.err
  echo "bench dyn=$dyn rc=$?"; cut -c1-400 gpurun_out/c84_bench_dyn${dyn}.json
  cp gpurun_out/c84_bench_dyn${dyn}.json gpurun_out/c84_bench_dyn${dyn}_$RANDOM.json
done
