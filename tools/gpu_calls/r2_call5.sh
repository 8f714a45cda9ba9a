#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
timeout 60 python tools/chain_timeline.py humanoid 4096 bf16x3 gelu 2>&1 | sed -n "/step 2/,\$p" | cut -c1-260 | head -24 > gpurun_out/chain_timeline_c5.txt
for rep in 1 2; do for v in 1 0; do
  DSACT_BWD_SPLIT=$v timeout 120 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c5_split${v}_$rep.json
done; done
timeout 300 python tools/e2e_diag.py > gpurun_out/e2e_diag_c5.txt 2>&1
