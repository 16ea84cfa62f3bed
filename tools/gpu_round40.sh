#!/bin/bash
mkdir -p gpurun_out
echo "=== build + pytest -m gpu ==="
timeout 900 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench N=1 (default flags) ==="
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-160
echo "=== bench --impl reference ==="
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-400
