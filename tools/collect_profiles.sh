#!/bin/bash
# One GPU box: everything profiles/ is refreshed from (bench lines, ncu launch lists, one full capture of the tcgen05 kernels,
# kernel timeline, batch sweep with clocks, CNN bench).  Outputs go to gpurun_out/; summarise here with tools/ncu_summary.py.
set -x
mkdir -p gpurun_out
python bench.py --steps 1000 --warmup 50 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
python bench.py --steps 500 --warmup 30 --gemm bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python bench.py --steps 200 --warmup 20 --gemm fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
DSACT_PDL=0 python tools/trace_step.py > gpurun_out/trace_step.txt 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bf16x3.csv python tools/ncu_target.py --steps 2 --gemm bf16x3 --replay-size 1000000 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_ -c 7 -f \
  -o gpurun_out/prof_tc_r2 python tools/ncu_target.py --steps 1 --gemm bf16x3 --replay-size 1000000 > gpurun_out/ncu2.log 2>&1
python tools/sweep.py --modes bf16x3,fp32 --batches 256,4096,65536 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
python tools/bench_cnn.py --cpu > gpurun_out/bench_cnn.json 2> gpurun_out/bench_cnn.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cnn.csv \
  python tools/bench_cnn.py --steps 1 --warmup 1 > gpurun_out/ncu3.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
