#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call 88 (1 GPU): everything at HEAD -- the whole GPU suite (incl. full-width parity), smoke(), the driver's two bench arms with
# default flags, C4 / C5 lines, a launch list of the C2 step and `ncu --set full` captures of the GEMM (cluster launch control) and
# attention kernels.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/c88_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/c88_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c88_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/c88_smoke.log
timeout 900 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/c88_bench_reference_arm.json 2> gpurun_out/c88_bench_reference_arm.err
echo "reference arm rc=$?"
timeout 900 python bench.py --steps 8 --warmup 3 --profile-step gpurun_out/c88_step_profile_c2.json > gpurun_out/c88_bench_c2.json 2> gpurun_out/c88_bench_c2.err
echo "bench c2 rc=$?"
timeout 900 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/c88_step_profile_c4.json > gpurun_out/c88_bench_c4.json 2> gpurun_out/c88_bench_c4.err
echo "bench c4 rc=$?"
timeout 900 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference --profile-step gpurun_out/c88_step_profile_c5.json > gpurun_out/c88_bench_c5.json 2> gpurun_out/c88_bench_c5.err
echo "bench c5 rc=$?"
python - <<'PY'
import json
for f in ("c88_bench_reference_arm", "c88_bench_c2", "c88_bench_c4", "c88_bench_c5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"], 1), d.get("ms_per_step"), d.get("clocks"), d.get("vs_gpu_reference"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1200:])
PY
# launch list of two C2 steps (L = 4 keeps it short; shares, not absolutes) and full captures of the hot kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c88_launches.csv \
    python bench.py --steps 1 --warmup 1 --layers 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/c88_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_|gemm_bf16' --launch-skip 7 -c 7 \
    -o gpurun_out/c88_kernels python tools/ncu_targets.py > gpurun_out/c88_ncu_targets.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/c88_kernels.ncu-rep 2>/dev/null
