#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call B (1 GPU): GPU suite with the new kernels (wgrad multi-launch + fp32 TMA epilogues, row-resident CE, TMA-staged
# LSE/Delta in the attention backward, MoE pair-mode grouped GEMM + gather4), micro-benchmarks, C2 / C4 bench, ncu captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullwidth.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "shape|passed|failed|Error|error" > gpurun_out/b_fullwidth.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_gpu_fullwidth.py --durations=8 > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
rm -f gpurun_out/b_probe.jsonl
for c in wgrad_multi_bench gemm_pair_bench ce_bench elementwise_bench_c2 attn_bench_c2 attn_bench_hd128 moe_layer_fwd_c4; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/b_probe.jsonl > /dev/null 2>&1
done
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/b_step_profile_c2.json > gpurun_out/b_bench_c2.json 2> gpurun_out/b_bench_c2.err
timeout 420 python bench.py --config c4 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/b_step_profile_c4.json > gpurun_out/b_bench_c4.json 2> gpurun_out/b_bench_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_|gemm_bf16|ce_rows' --launch-skip 8 -c 8 -f -o gpurun_out/r02_kernels_hd80 python tools/ncu_targets.py > gpurun_out/b_ncu.log 2>&1
cat gpurun_out/b_fullwidth.log
tail -c 1500 gpurun_out/b_pytest.log
python - <<'PY'
import json
for l in open("gpurun_out/b_probe.jsonl"):
    d = json.loads(l); print(json.dumps(d)[:1400])
for c in ("c2", "c4"):
    try:
        d = json.load(open(f"gpurun_out/b_bench_{c}.json")); print(c, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["peak_hbm_gb"])
    except Exception as e:
        print(c, "bench failed", e); print(open(f"gpurun_out/b_bench_{c}.err").read()[-1500:])
PY
tail -5 gpurun_out/b_ncu.log
