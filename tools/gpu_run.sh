#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 94 (1 GPU, the round's last GPU minutes): compute-sanitizer memcheck over the kernel-level GPU tests of HEAD (round-2 kernels:
# cluster-launch-control GEMM incl. the zero-writing K-grouped epilogue, split-softmax attention, dropout, decode, MoE), then the
# C5 line at HEAD if time is left.
set -u
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_moe.py \
    tests/test_gpu_dropout.py tests/test_zzz_generation.py -q -m gpu -k "not full_size" -p no:cacheprovider \
    > gpurun_out/c94_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/c94_memcheck.log | tail -n 8 | cut -c1-300
timeout 150 python bench.py --config c5 --no-cpu-baseline --no-gpu-reference > gpurun_out/c94_bench_c5.json 2> gpurun_out/c94_bench_c5.err
echo "bench c5 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c94_bench_c5.json") if l.startswith("{")][-1])
    print("c5", round(d["value"], 1), round(d.get("ms_per_step"), 2), d.get("clocks"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
except Exception as e:
    print("c5 failed", e); print(open("gpurun_out/c94_bench_c5.err").read()[-800:])
PY
