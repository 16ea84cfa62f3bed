#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 93 (1 GPU): HEAD with the lazily cleared MoE expert gradients -- whole GPU suite, smoke(), C4 with and without the lazy
# clearing on the same box, C2 line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/c93_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/c93_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c93_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/c93_smoke.log
for mode in eager lazy; do  # the first run gets the cooler chip: a win of the second is conservative
  case $mode in eager*) export DOLO_EAGER_GRAD_ZERO=1;; *) export DOLO_EAGER_GRAD_ZERO=0;; esac
  extra=""; [ "$mode" = lazy ] && extra="--profile-step gpurun_out/c93_step_profile_c4.json"
  timeout 240 python bench.py --config c4 --no-cpu-baseline --no-gpu-reference $extra > gpurun_out/c93_bench_c4_$mode.json 2> gpurun_out/c93_bench_c4_$mode.err
  echo "bench c4 $mode rc=$?"
done
unset DOLO_EAGER_GRAD_ZERO
timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference > gpurun_out/c93_bench_c2.json 2> gpurun_out/c93_bench_c2.err
echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ("c93_bench_c4_eager", "c93_bench_c4_lazy", "c93_bench_c2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 1), round(d.get("ms_per_step"), 2), d.get("clocks"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1200:])
PY
