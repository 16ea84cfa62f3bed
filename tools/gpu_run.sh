#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call H (1 GPU): KV-cache decoding tests, the whole suite, smoke(), and the driver's two bench arms with default flags.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_generation.py -m gpu -q > gpurun_out/h_generation.log 2>&1
echo "rc=$?" >> gpurun_out/h_generation.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_fullwidth.py > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/h_smoke.log
timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/h_bench_reference_arm.json 2> gpurun_out/h_bench_reference_arm.err
timeout 900 python bench.py --steps 8 --warmup 3 --profile-step gpurun_out/h_step_profile_c2.json > gpurun_out/h_bench_c2.json 2> gpurun_out/h_bench_c2.err
tail -c 900 gpurun_out/h_generation.log
tail -c 400 gpurun_out/h_pytest.log
cat gpurun_out/h_smoke.log | tail -3
python - <<'PY'
import json
for f in ("h_bench_reference_arm", "h_bench_c2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d.get("ms_per_step"), d.get("clocks"), d.get("vs_gpu_reference"), (d.get("cpu_baseline") or {}).get("sample", "")[:160], d.get("wall_s"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
