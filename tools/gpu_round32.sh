#!/bin/bash
# final check of the round: new generation tests first (fast signal), then the whole GPU suite
mkdir -p gpurun_out
echo "=== generation tests ==="
timeout 400 python -m pytest tests/test_generation.py -q -m gpu 2>&1 | tail -25
echo "=== pytest -m gpu (all) ==="
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
