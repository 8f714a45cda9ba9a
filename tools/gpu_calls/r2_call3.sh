#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
timeout 60 python tools/chain_timeline.py 2>&1 | sed -n "/step 2/,\$p" | cut -c1-330 | head -14 > gpurun_out/chain_timeline_packed.txt
timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_packed.json
