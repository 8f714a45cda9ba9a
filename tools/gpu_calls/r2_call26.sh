#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs" | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1000 --warmup 50 > gpurun_out/bench_dp2_final.json 2> gpurun_out/bench_dp2_final.err
python -c "import json; d=json.load(open('gpurun_out/bench_dp2_final.json')); print(d['value'], d['ms_per_step'], d['dp_check'], d['e2e']['value'])"
