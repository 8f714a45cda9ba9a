#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_std.py -m gpu -x -q 2>&1 | tail -12
