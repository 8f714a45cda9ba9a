#!/bin/bash
# streamed chain kernel: first hardware run (watchdog build), parity, timelines, A/B bench
set -x
mkdir -p gpurun_out
G=$PWD/dsac-v2_b200/libdsact_guard.so
DSACT_LIB=$G timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
DSACT_LIB=$G timeout 60 python tools/chain_timeline.py 2>&1 | sed -n "/step 2/,\$p" | cut -c1-330 | head -14 > gpurun_out/chain_timeline_stream.txt
timeout 300 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
for v in 1 0; do
  DSACT_CHAIN_STREAM=$v timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_stream_$v.json
done
DSACT_PDL=0 timeout 120 python tools/trace_step.py > gpurun_out/trace_step_stream.txt 2>/dev/null
