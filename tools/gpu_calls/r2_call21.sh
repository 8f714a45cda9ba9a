#!/bin/bash
set -x
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs\|return float" | tail -8
python bench.py --steps 1000 --warmup 50 --no-cpu-baseline > gpurun_out/bench_c21.json 2> gpurun_out/bench_c21.err
cat gpurun_out/bench_c21.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
