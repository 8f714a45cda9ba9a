#!/bin/bash
set -x
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs\|return float" | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
for v in prev new prev new; do
if [ $v = prev ]; then export DSACT_LIB=$PWD/tools/ab/libdsact_prev.so; else unset DSACT_LIB; fi
python bench.py --steps 1500 --warmup 50 --no-cpu-baseline > gpurun_out/bench_c28_$v.json 2> gpurun_out/bench_c28.err
python -c "import json; d=json.load(open('gpurun_out/bench_c28_$v.json')); print('$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
unset DSACT_LIB
DSACT_PDL=0 python tools/trace_step.py > gpurun_out/trace_step_c28.txt 2>/dev/null; tail -4 gpurun_out/trace_step_c28.txt
