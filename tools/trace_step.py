#!/usr/bin/env python
"""Kernel timeline of one captured step (CUPTI through torch.profiler; no serialisation, unlike ncu):
start offset, duration and stream of every kernel of the step, plus the idle time between them.

    python tools/trace_step.py [--gemm bf16x3] [--batch 4096]
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from dsac_v2_b200 import synth  # noqa: E402
from dsac_v2_b200.engine import Engine, make_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--replay-size", type=int, default=200_000)
ap.add_argument("--gemm", default="bf16x3")
ap.add_argument("--steps", type=int, default=6)
a = ap.parse_args()

cfg = synth.CONFIGS["humanoid"]
B = a.batch
lim = torch.full((17,), 0.4)
eng = Engine(make_config(376, 17, cfg["hidden"], cfg["hidden"], max_batch=B, gemm_mode=a.gemm), torch.device("cuda", 0), lim, -lim)
eng.load_weights(synth.make_weights(cfg))
eng.bind_replay(a.replay_size)
g = torch.Generator(device="cuda").manual_seed(123)
r = eng.replay
r["obs"].normal_(generator=g); r["obs2"].normal_(generator=g); r["rew"].normal_(generator=g)
r["act"].uniform_(-0.4, 0.4, generator=g)
it = 0
for _ in range(20):
    eng.replay_step(B, a.replay_size, it); it += 1
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        eng.replay_step(B, a.replay_size, it); it += 1
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
names = [e.name for e in ev]
first = names[0]
starts = [i for i, n in enumerate(names) if n == first]
per = len(ev) // a.steps
lo = per * (a.steps - 2)          # second-to-last step
step = ev[lo:lo + per]
t0 = step[0].time_range.start
nxt = ev[lo + per].time_range.start if lo + per < len(ev) else None
print(f"# {per} device activities per step; step span {(nxt - t0) if nxt else float('nan'):.1f} us")
busy_end = t0
for e in step:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - busy_end
    busy_end = max(busy_end, e.time_range.end)
    print(f"{s:8.1f} us  +{d:6.1f} us  gap {gap:6.1f}  {e.name[:70]}")
