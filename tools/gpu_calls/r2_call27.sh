#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs" | tail -4
DSACT_DP_TWO_SHOT=1 timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | grep -v "Warning\|^$\|Docs" | tail -4
for cfg in "1 0" "0 0" "1 1" "0 1" "1 0" "0 0"; do
set -- $cfg
DSACT_DP_SPLIT=$1 DSACT_DP_TWO_SHOT=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1500 --warmup 50 --no-cpu-baseline > gpurun_out/bench_dp2_s$1_t$2.json 2> gpurun_out/bench_dp2_ab.err
python -c "import json; d=json.load(open('gpurun_out/bench_dp2_s$1_t$2.json')); print('split$1 twoshot$2', d['value'], d['ms_per_step'], d['dp_check']['status'], d['dp_check']['replicas_bit_identical'], d['gpu_launches'])"
done
