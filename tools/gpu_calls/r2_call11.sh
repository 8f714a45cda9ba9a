#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cnn.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/ -m gpu -x -q --deselect tests/test_gpu_cnn.py 2>&1 | tail -4
timeout 200 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_c11.json
python -c "import json;d=json.load(open('gpurun_out/bench_c11.json'));print(round(d['value'],1),round(d['ms_per_step'],5),d['launches_per_step'],round(d['e2e']['value'],1),d['e2e']['h2d_gbs'],round(d['roofline']['other_ms_per_step'],4),{k:round(x['ms_per_step'],4) for k,x in d['roofline']['by_kind'].items()})"
