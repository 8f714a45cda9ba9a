#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -4
for v in 1 0 1 0; do
  DSACT_JOBS=$v timeout 120 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c9_jobs${v}.json
  python -c "import json;d=json.load(open('gpurun_out/bench_c9_jobs${v}.json'));print('jobs=$v',d['value'],d['ms_per_step'],d['launches_per_step'],d['e2e']['value'],d['e2e']['h2d_gbs'],d['roofline']['other_ms_per_step'],{k:round(x['ms_per_step'],4) for k,x in d['roofline']['by_kind'].items()})"
done
DSACT_PDL=0 timeout 120 python tools/trace_step.py > gpurun_out/trace_step_c9.txt 2>/dev/null
timeout 60 python tools/chain_timeline.py humanoid 4096 bf16x3 gelu 2>&1 | sed -n "/step 2/,\$p" | cut -c1-260 | head -50 > gpurun_out/chain_timeline_c9.txt
