#!/bin/bash
# 8 GPUs: data-parallel equivalence at world 4 and 8, scaling lines at N = 8 and 4, one-shot vs two-shot exchange, config 3
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -k "peer-4 or peer-8 or nccl-4 or nccl-8" 2>&1 | grep -v "Warning\|warn\|return float\|^$\|Docs" | tail -8
run() { # name N extra-args env...
  name=$1; N=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 500 --warmup 30 $EXTRA > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print('$name',d['n_gpus'],round(d['value'],1),round(d['ms_per_step'],5),d['config']['parallelism'],round(d['e2e']['value'],1),(d['dp_check'] or {}).get('status'),d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_$name.err
}
EXTRA="" run dp8_twoshot 8 DSACT_DP_TWO_SHOT=1
EXTRA="" run dp8_oneshot 8 DSACT_DP_TWO_SHOT=0
EXTRA="" run dp4_twoshot 4 DSACT_DP_TWO_SHOT=1 CUDA_VISIBLE_DEVICES=0,1,2,3
EXTRA="" run dp4_oneshot 4 DSACT_DP_TWO_SHOT=0 CUDA_VISIBLE_DEVICES=0,1,2,3
EXTRA="--config halfcheetah --batch 8192" run cfg3_halfcheetah_dp8 8 DSACT_DP_TWO_SHOT=1
EXTRA="--dp nccl" run dp8_nccl 8 DSACT_DP_TWO_SHOT=1
