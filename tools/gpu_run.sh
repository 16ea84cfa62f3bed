#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call F (1 GPU): attention forward with an exact alpha == 1 when the reference maximum stays (both kernels), pipelined backward
# with bank-conflict-free swizzled dQ slabs + TMA tile reduce, RoPE with two tokens in flight; whole suite; three bench lines.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_fullwidth.py > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
rm -f gpurun_out/f_probe.jsonl
for c in attn_bench_c2 attn_bench_hd128 elementwise_bench_c2; do timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/f_probe.jsonl > /dev/null 2>&1; done
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/f_step_profile_c2.json > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err
timeout 420 python bench.py --config c4 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/f_step_profile_c4.json > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err
timeout 420 python bench.py --config c5 --steps 3 --warmup 3 --checkpoint-every 1 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/f_step_profile_c5.json > gpurun_out/f_bench_c5.json 2> gpurun_out/f_bench_c5.err
timeout 900 python -m pytest tests/test_gpu_fullwidth.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "shape:|passed|failed|Error|error" > gpurun_out/f_fullwidth.log
tail -c 700 gpurun_out/f_pytest.log
python - <<'PY'
import json
for l in open("gpurun_out/f_probe.jsonl"):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k not in ("vs_flash_dk", "vs_flash_dv", "trace", "vs_flash_fwd", "single_buffer_vs_flash_fwd")})
for c in ("c2", "c4", "c5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/f_bench_{c}.json") if l.startswith("{")][-1]); print(c, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["peak_hbm_gb"], d["clocks"])
    except Exception as e:
        print(c, "bench failed", e); print(open(f"gpurun_out/f_bench_{c}.err").read()[-1500:])
PY
cut -c1-260 gpurun_out/f_fullwidth.log
