#!/bin/bash
mkdir -p gpurun_out
echo "=== CTA-pair GEMM ==="
python tools/gpu_probe.py --only gemm_pair --timeout 200 2>&1 | tail -4 | cut -c1-1500
python - <<'PY'
import json
for l in open('gpurun_out/probe.jsonl'):
    r = json.loads(l)
    if r['case'].startswith('gemm_pair'):
        print(json.dumps(r)[:1800])
PY
echo "=== pytest gpu ==="
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench (panel raster, attn v3) ==="
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r9.json | cut -c1-300
