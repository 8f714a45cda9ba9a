#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"apply_kernel|sample_kernel|loss_kernel|policy_grad_kernel|gather_kernel|step_prologue" -c 6 -f \
  -o gpurun_out/r2_small python tools/ncu_target.py --steps 1 --gemm bf16x3 --replay-size 1000000 > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm -c 3 -f \
  -o gpurun_out/r2_wgrad python tools/ncu_target.py --steps 1 --gemm bf16x3 > gpurun_out/ncu_wgrad.log 2>&1
tail -2 gpurun_out/ncu_wgrad.log
timeout 300 python tools/e2e_diag.py > gpurun_out/e2e_diag_c6.txt 2>&1
