#!/bin/bash
set -x
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs\|return float" | tail -6
for v in 0 1 0 1; do
DSACT_APPLY_SPLIT=$v python bench.py --steps 1500 --warmup 50 --no-cpu-baseline > gpurun_out/bench_c23_split$v.json 2> gpurun_out/bench_c23.err
python -c "import sys,json; d=json.load(open('gpurun_out/bench_c23_split$v.json')); print('split$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
DSACT_PDL=0 python tools/trace_step.py > gpurun_out/trace_step_c23.txt 2>/dev/null; cat gpurun_out/trace_step_c23.txt
