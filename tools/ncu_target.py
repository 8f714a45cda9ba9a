#!/usr/bin/env python
"""Profiling target: a few device-resident DSAC-T steps (bench.py's workload) bracketed by
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (B200_PROFILING.md commands).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_target.py --steps 2
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from dsac_v2_b200 import synth  # noqa: E402
from dsac_v2_b200.engine import Engine, make_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--replay-size", type=int, default=200_000)
ap.add_argument("--gemm", default="fp32")
ap.add_argument("--eager", action="store_true", help="no CUDA graph (plain stream launches)")
a = ap.parse_args()

cfg = synth.CONFIGS["humanoid"]
B = a.batch
lim = torch.full((17,), 0.4)
eng = Engine(make_config(376, 17, cfg["hidden"], cfg["hidden"], max_batch=B, gemm_mode=a.gemm, use_graph=not a.eager),
             torch.device("cuda", 0), lim, -lim)
eng.load_weights(synth.make_weights(cfg))
eng.bind_replay(a.replay_size)
g = torch.Generator(device="cuda").manual_seed(123)
r = eng.replay
r["obs"].normal_(generator=g); r["obs2"].normal_(generator=g); r["rew"].normal_(generator=g)
r["act"].uniform_(-0.4, 0.4, generator=g)
it = 0
for _ in range(3):
    eng.replay_step(B, a.replay_size, it); it += 1
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    eng.replay_step(B, a.replay_size, it); it += 1
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per step:", eng.last_call_launches())
