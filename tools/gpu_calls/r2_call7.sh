#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 500 --warmup 30 > gpurun_out/bench_c7.json 2> gpurun_out/bench_c7.err
tail -3 gpurun_out/bench_c7.err
DSACT_PDL=0 timeout 120 python tools/trace_step.py > gpurun_out/trace_step_c7.txt 2>/dev/null
