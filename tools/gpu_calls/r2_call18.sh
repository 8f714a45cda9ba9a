#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_cnn.py -m gpu -x -q 2>&1 | tail -5
python tools/bench_cnn.py > gpurun_out/bench_cnn_c18.json 2> gpurun_out/bench_cnn_c18.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cnn_c18.csv \
  python tools/bench_cnn.py --steps 1 --warmup 1 > gpurun_out/ncu3.log 2>&1
cat gpurun_out/bench_cnn_c18.json
