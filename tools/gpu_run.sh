#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call C (1 GPU): the whole GPU suite on HEAD, micro-benchmarks (wgrad epilogue / multi-launch variants, L2 hints A/B, CE,
# attention with the lazy rescale, MoE layer), C2 and C4 bench lines, ncu capture of the same targets.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullwidth.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "shape|passed|failed|Error|error" > gpurun_out/c_fullwidth.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_fullwidth.py --durations=8 > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
rm -f gpurun_out/c_probe.jsonl
for c in wgrad_multi_bench gemm_pair_bench ce_bench attn_bench_c2 attn_bench_hd128 moe_layer_fwd_c4; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/c_probe.jsonl > /dev/null 2>&1
done
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/c_step_profile_c2.json > gpurun_out/c_bench_c2.json 2> gpurun_out/c_bench_c2.err
timeout 420 python bench.py --config c4 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/c_step_profile_c4.json > gpurun_out/c_bench_c4.json 2> gpurun_out/c_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_|gemm_bf16|ce_rows' --launch-skip 8 -c 8 -f -o gpurun_out/r02_kernels_hd80_c python tools/ncu_targets.py > gpurun_out/c_ncu.log 2>&1
cat gpurun_out/c_fullwidth.log
tail -c 1800 gpurun_out/c_pytest.log
python - <<'PY'
import json
for l in open("gpurun_out/c_probe.jsonl"):
    d = json.loads(l); print(json.dumps(d)[:1700])
for c in ("c2", "c4"):
    try:
        d = json.load(open(f"gpurun_out/c_bench_{c}.json")); print(c, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["peak_hbm_gb"], d["clocks"])
    except Exception as e:
        print(c, "bench failed", e); print(open(f"gpurun_out/c_bench_{c}.err").read()[-1500:])
PY
tail -3 gpurun_out/c_ncu.log
