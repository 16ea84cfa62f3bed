#!/bin/bash
# memory-safety pass over the final kernels: compute-sanitizer memcheck on the kernel test suite (small shapes)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 0 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu \
  -k "not full_size" 2>&1 | tail -15 | tee gpurun_out/memcheck_kernels.txt
echo "memcheck rc=${PIPESTATUS[0]}"
