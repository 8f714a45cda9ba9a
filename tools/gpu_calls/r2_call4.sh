#!/bin/bash
set -x
mkdir -p gpurun_out
cd tools/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate > ../../gpurun_out/fma_rate.txt 2>&1; cd ../..
timeout 400 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
timeout 60 python tools/chain_timeline.py humanoid 4096 bf16x3 relu 2>&1 | sed -n "/step 2/,\$p" | cut -c1-260 | head -8 > gpurun_out/chain_timeline_relu.txt
timeout 60 python tools/chain_timeline.py humanoid 4096 bf16 gelu 2>&1 | sed -n "/step 2/,\$p" | cut -c1-260 | head -8 > gpurun_out/chain_timeline_bf16.txt
timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c4.json
DSACT_FOLD_TAIL=0 DSACT_PROLOGUE_MERGE=0 timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c4_nofold.json
DSACT_PDL=0 timeout 120 python tools/trace_step.py > gpurun_out/trace_step_c4.txt 2>/dev/null
timeout 120 python tools/trace_step.py > gpurun_out/trace_step_c4_pdl.txt 2>/dev/null
