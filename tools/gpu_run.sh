#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call A (1 GPU): full GPU suite incl. the full-width parity tests, then the three workloads with the gpu_reference arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/a_env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 -x --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 --profile-step gpurun_out/a_step_profile_c2.json > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err
timeout 420 python bench.py --config c4 --steps 4 --warmup 3 --no-cpu-baseline --profile-step gpurun_out/a_step_profile_c4.json > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err
timeout 420 python bench.py --config c5 --steps 3 --warmup 3 --checkpoint-every 1 --no-cpu-baseline --profile-step gpurun_out/a_step_profile_c5.json > gpurun_out/a_bench_c5.json 2> gpurun_out/a_bench_c5.err
tail -c 600 gpurun_out/a_pytest.log
for f in gpurun_out/a_bench_c*.json; do echo "== $f"; head -c 1500 $f; echo; done
for f in gpurun_out/a_bench_c*.err; do echo "== $f"; tail -c 800 $f; done
