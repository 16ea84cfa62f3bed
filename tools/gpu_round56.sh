#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k sft_feed 2>&1 | grep -E "Error|error|assert|^E " | head -30
