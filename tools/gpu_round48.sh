#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== bench A/B (fused SwiGLU epilogue on / off), same box ==="
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fused.json | cut -c1-150
DOLO_NO_SWIGLU_FUSION=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_unfused.json | cut -c1-150
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fused2.json | cut -c1-150
