#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call E (1 GPU): the split-softmax attention forward (double-buffered S, two threads per row): tests, A/B against the
# single-buffer kernel and flash-attn at hd 80 / 128, C2 and C5 bench lines, ncu capture.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > gpurun_out/e_attn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/e_attn_tests.log
rm -f gpurun_out/e_probe.jsonl
for c in attn_bench_c2 attn_bench_hd128; do timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/e_probe.jsonl > /dev/null 2>&1; done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullwidth.py -m gpu -q -x -s -p no:cacheprovider 2>&1 | grep -E "shape|passed|failed|Error|error" > gpurun_out/e_model_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/e_step_profile_c2.json > gpurun_out/e_bench_c2.json 2> gpurun_out/e_bench_c2.err
timeout 420 python bench.py --config c5 --steps 3 --warmup 3 --checkpoint-every 1 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/e_step_profile_c5.json > gpurun_out/e_bench_c5.json 2> gpurun_out/e_bench_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd' --launch-skip 1 -c 1 -f -o gpurun_out/r02_attn_fwd_split_hd80 python tools/ncu_targets.py > gpurun_out/e_ncu.log 2>&1
tail -c 600 gpurun_out/e_attn_tests.log
python - <<'PY'
import json
for l in open("gpurun_out/e_probe.jsonl"):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k not in ("vs_flash_dk", "vs_flash_dv", "trace")})
for c in ("c2", "c5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/e_bench_{c}.json") if l.startswith("{")][-1]); print(c, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["peak_hbm_gb"], d["clocks"])
    except Exception as e:
        print(c, "bench failed", e); print(open(f"gpurun_out/e_bench_{c}.err").read()[-1500:])
PY
cat gpurun_out/e_model_tests.log | cut -c1-300
tail -3 gpurun_out/e_ncu.log
