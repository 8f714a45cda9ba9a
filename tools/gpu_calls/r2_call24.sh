#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_v1.py tests/test_gpu_std.py tests/test_gpu_cnn.py -m gpu -q 2>&1 | tail -30
