#!/bin/bash
set -x
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_v1.py -m gpu -x -q -k "oracle" > gpurun_out/racecheck_heads.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck_heads.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_update_matches_reference_golden and (tiny_b16 or ragged_b37 or tiny_gauss)" > gpurun_out/memcheck_fp32.log 2>&1
echo "memcheck fp32 rc=$?"; tail -6 gpurun_out/memcheck_fp32.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_bf16x3_tensor_core_path_matches_reference_golden and tiny_b16" > gpurun_out/memcheck_tc.log 2>&1
echo "memcheck tcgen05 rc=$?"; tail -8 gpurun_out/memcheck_tc.log
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
