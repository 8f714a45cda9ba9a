#!/bin/bash
# N GPUs (gpurun --gpus N): data-parallel equivalence tests for every world size that fits, then the scaling bench at N
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_cnn.py -m gpu -x -q 2>&1 | grep -v "Warning\|warn\|return float\|^$\|Docs" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 30 > gpurun_out/bench_dp${N}.json 2> gpurun_out/bench_dp${N}.err
tail -3 gpurun_out/bench_dp${N}.err
python -c "import json;d=json.load(open('gpurun_out/bench_dp${N}.json'));print(d['n_gpus'],round(d['value'],1),round(d['ms_per_step'],5),d['config']['parallelism'],round(d['e2e']['value'],1),d['dp_check'])"
