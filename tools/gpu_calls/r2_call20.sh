#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_std.py tests/test_gpu_cnn.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs\|return float" | tail -6
timeout 300 python tools/trainer_rate.py > gpurun_out/trainer_rate.jsonl 2> gpurun_out/trainer_rate.err
timeout 300 python tools/trainer_rate.py --batch 4096 >> gpurun_out/trainer_rate.jsonl 2>> gpurun_out/trainer_rate.err
cat gpurun_out/trainer_rate.jsonl
