#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_dropin.py tests/test_gpu_loop.py -m gpu -x -q 2>&1 | tail -5
python tools/bench_cnn.py > gpurun_out/bench_cnn_c19.json 2> gpurun_out/bench_cnn_c19.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cnn_c19.csv \
  python tools/bench_cnn.py --steps 1 --warmup 1 > gpurun_out/ncu3.log 2>&1
cat gpurun_out/bench_cnn_c19.json
timeout 600 python tools/trainer_rate.py > gpurun_out/trainer_rate.jsonl 2> gpurun_out/trainer_rate.err
timeout 600 python tools/trainer_rate.py --batch 4096 >> gpurun_out/trainer_rate.jsonl 2>> gpurun_out/trainer_rate.err
cat gpurun_out/trainer_rate.jsonl; tail -5 gpurun_out/trainer_rate.err
