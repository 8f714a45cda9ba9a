#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q 2>&1 | grep -v "Warning\|warn\|return float\|^$\|Docs" | tail -6
timeout 300 python bench.py --steps 1000 --warmup 50 > gpurun_out/bench_c12.json 2> gpurun_out/bench_c12.err
python -c "import json;d=json.load(open('gpurun_out/bench_c12.json'));print(round(d['value'],1),round(d['ms_per_step'],5),d['launches_per_step'],d['e2e'],round(d['roofline']['other_ms_per_step'],4),{k:round(x['ms_per_step'],4) for k,x in d['roofline']['by_kind'].items()}, d['cuda_eager_baseline']['value'], d['cpu_baseline']['value'])"
timeout 300 python tools/bench_cnn.py --cpu > gpurun_out/bench_cnn_c12.json 2> gpurun_out/bench_cnn_c12.err; cat gpurun_out/bench_cnn_c12.json; tail -2 gpurun_out/bench_cnn_c12.err
timeout 60 python tools/chain_timeline.py humanoid 4096 bf16x3 gelu 2>&1 | sed -n "/step 2/,\$p" | cut -c1-260 | head -12 > gpurun_out/chain_timeline_c12.txt
