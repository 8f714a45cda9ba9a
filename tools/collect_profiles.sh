#!/bin/bash
# One GPU box: everything profiles/ is refreshed from (bench lines, ncu launch list, one full capture, batch sweep).
# Outputs go to gpurun_out/; summarise here with tools/ncu_summary.py and copy into profiles/.
set -x
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
python bench.py --steps 200 --warmup 20 --gemm bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python bench.py --steps 100 --warmup 10 --gemm fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
DSACT_PDL=0 python tools/trace_step.py > gpurun_out/trace_step.txt 2>/dev/null
python tools/chain_timeline.py 2>&1 | sed -n "/step 2/,\$p" | cut -c1-400 > gpurun_out/chain_timeline.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bf16x3.csv python tools/ncu_target.py --steps 2 --gemm bf16x3 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_ -c 7 -f \
  -o gpurun_out/prof_tc_final python tools/ncu_target.py --steps 1 --gemm bf16x3 > gpurun_out/ncu2.log 2>&1
python tools/sweep.py --modes bf16x3,bf16 --batches 256,1024,4096,16384,65536 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log
