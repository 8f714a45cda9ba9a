#!/bin/bash
set -x
# memcheck of the head-wise fp32 engine (conv kernels, heads GEMMs, V1 loss): no tcgen05 / TMA kernels in these tests
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_v1.py tests/test_gpu_std.py -m gpu -x -q -k "oracle or golden" > gpurun_out/memcheck_heads.log 2>&1
echo "memcheck rc=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_heads.log; tail -15 gpurun_out/memcheck_heads.log
