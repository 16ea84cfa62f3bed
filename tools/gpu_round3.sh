#!/bin/bash
# GPU call 3: parity suite (non-hd64 first), hd64 diagnosis, smoke, first bench
mkdir -p gpurun_out
echo "=== pytest gpu (without head_dim 64 cases) ==="
timeout 900 python -m pytest tests -m gpu -q -k "not c1 and not 64" -p no:cacheprovider 2>&1 | tail -40
echo "=== hd64 diagnosis ==="
for m in fwd bwd; do
  echo "--- $m hd64 [128]"; timeout 60 python tools/diag_attn.py $m 64 2 1 128 2>&1 | tail -5; echo "rc=$?"
done
echo "--- fwd hd64 [256]"; timeout 60 python tools/diag_attn.py fwd 64 2 1 256 2>&1 | tail -5; echo "rc=$?"
echo "--- bwd hd128 [128] (control)"; timeout 60 python tools/diag_attn.py bwd 128 2 1 128 2>&1 | tail -5; echo "rc=$?"
echo "--- compute-sanitizer fwd hd64"; timeout 200 compute-sanitizer --tool memcheck python tools/diag_attn.py fwd 64 2 1 128 2>&1 | tail -25
echo "=== smoke ==="
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== bench L=4 ==="
timeout 600 python bench.py --layers 4 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
echo "=== bench full ==="
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_first.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
