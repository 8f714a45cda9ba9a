#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -rs 2>&1 | grep -v "Warning\|^$\|Docs" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 1000 --warmup 50 --no-cpu-baseline > gpurun_out/bench_dp4_final.json 2> gpurun_out/bench_dp4_final.err
python -c "import json; d=json.load(open('gpurun_out/bench_dp4_final.json')); print(d['value'], d['ms_per_step'], d['dp_check']['status'], d['dp_check']['replicas_bit_identical'], d['e2e']['value'])"
