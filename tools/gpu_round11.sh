#!/bin/bash
mkdir -p gpurun_out
echo "=== attention fwd v2 (ping-pong) ==="
python tools/gpu_probe.py --only attn_fwd_v2 --timeout 150 2>&1 | tail -2 | cut -c1-300
python tools/gpu_probe.py --only attn_bench_c2_fwd_v2 --timeout 150 2>&1 | tail -1 | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/probe.jsonl'):
    r = json.loads(l)
    if 'fwd_v2' in r['case']:
        print(json.dumps({k: v for k, v in r.items() if not k.startswith('vs_')})[:1500])
PY
echo "=== ncu launch list mbs4 L2 ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_L2_mbs4.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_L2_mbs4.csv
