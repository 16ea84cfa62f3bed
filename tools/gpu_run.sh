#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 92 (1 GPU): final state -- the whole GPU suite, smoke(), and the driver's two bench arms with default flags.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/c92_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/c92_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c92_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/c92_smoke.log
timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/c92_bench_reference_arm.json 2> gpurun_out/c92_bench_reference_arm.err
echo "reference arm rc=$?"
timeout 900 python bench.py > gpurun_out/c92_bench_c2_default_flags.json 2> gpurun_out/c92_bench_c2.err
echo "bench c2 (no flags) rc=$?"
python - <<'PY'
import json
for f in ("c92_bench_reference_arm", "c92_bench_c2_default_flags"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 1), d.get("ms_per_step"), d.get("clocks"), d.get("vs_gpu_reference"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1200:])
PY
