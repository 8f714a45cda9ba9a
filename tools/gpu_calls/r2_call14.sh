#!/bin/bash
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
PYTHONFAULTHANDLER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 30 > gpurun_out/bench_dp${N}.json 2> gpurun_out/bench_dp${N}.err
echo "rc=$?"
tail -40 gpurun_out/bench_dp${N}.err
cat gpurun_out/bench_dp${N}.json | head -c 3000
