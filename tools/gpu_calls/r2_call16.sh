#!/bin/bash
# 2 GPUs: why does bench.py under torchrun print nothing after a pytest run?
set -x
mkdir -p gpurun_out
N=2
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -k "bf16x3-peer-2" 2>&1 | tail -3
ps aux | grep -c python
nvidia-smi --query-compute-apps=pid,used_memory --format=csv
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 --tee 3 --log-dir gpurun_out/tlog bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/dbg_bench.out 2> gpurun_out/dbg_bench.err
echo "rc=$?"
wc -c gpurun_out/dbg_bench.out gpurun_out/dbg_bench.err
tail -20 gpurun_out/dbg_bench.err | cut -c1-400
find gpurun_out/tlog -type f | head; for f in $(find gpurun_out/tlog -name "*.log" | head -4); do echo "== $f"; tail -15 $f | cut -c1-300; done
