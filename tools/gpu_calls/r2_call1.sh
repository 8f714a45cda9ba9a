#!/bin/bash
# Round-2 first GPU call: baseline sanity, first hardware run of the split chain kernel (watchdog build), source-level
# ncu capture of the chain kernels, end-to-end host-path diagnostics.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16x3" 2>&1 | tail -3
# split variant, watchdog build
DSACT_LIB=$PWD/dsac-v2_b200/libdsact_guard.so DSACT_CHAIN_SPLIT=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16x3" 2>&1 | tail -5
DSACT_LIB=$PWD/dsac-v2_b200/libdsact_guard.so DSACT_CHAIN_SPLIT=1 timeout 60 python tools/chain_timeline.py 2>&1 | sed -n "/step 2/,\$p" | cut -c1-330 | head -8 > gpurun_out/chain_timeline_split.txt
timeout 60 python tools/chain_timeline.py 2>&1 | sed -n "/step 2/,\$p" | cut -c1-330 | head -12 > gpurun_out/chain_timeline_base.txt
for v in 0 1; do
  DSACT_CHAIN_SPLIT=$v timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_split_$v.json
done
DSACT_PDL=0 timeout 120 python tools/trace_step.py > gpurun_out/trace_step_base.txt 2>/dev/null
timeout 300 python tools/e2e_diag.py > gpurun_out/e2e_diag.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_chain -c 4 -f \
  -o gpurun_out/r2_chain_src python tools/ncu_target.py --steps 1 --gemm bf16x3 > gpurun_out/ncu_chain.log 2>&1
tail -2 gpurun_out/ncu_chain.log
ls -la gpurun_out
