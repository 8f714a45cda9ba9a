#!/usr/bin/env python
"""Batch-size sweep of the device-resident DSAC-T step (BASELINE.json config "S": synthetic obs=376 act=17,
256 -> 131072), one GPU: steps/s, samples/s and algorithmic TFLOP/s per arithmetic mode."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks / throttle reasons during the timed region)
from dsac_v2_b200 import synth  # noqa: E402
from dsac_v2_b200.engine import Engine, make_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="256,1024,4096,16384,65536,131072")
ap.add_argument("--modes", default="bf16x3,bf16,fp32")
ap.add_argument("--replay-size", type=int, default=1_000_000)
a = ap.parse_args()
cfg = synth.CONFIGS["humanoid"]
FLOP = 2 * 3_240_448
lim = torch.full((17,), 0.4)
out = []
for mode in a.modes.split(","):
    for B in [int(x) for x in a.batches.split(",")]:
        eng = Engine(make_config(376, 17, cfg["hidden"], cfg["hidden"], max_batch=B, gemm_mode=mode), torch.device("cuda", 0), lim, -lim)
        eng.load_weights(synth.make_weights(cfg))
        eng.bind_replay(a.replay_size)
        g = torch.Generator(device="cuda").manual_seed(123)
        r = eng.replay
        r["obs"].normal_(generator=g); r["obs2"].normal_(generator=g); r["rew"].normal_(generator=g)
        r["act"].uniform_(-0.4, 0.4, generator=g)
        it = 0
        for _ in range(5):
            eng.replay_step(B, a.replay_size, it); it += 1
        torch.cuda.synchronize()
        est = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        est[0].record()
        eng.replay_step(B, a.replay_size, it); it += 1
        est[1].record(); torch.cuda.synchronize()
        n = max(20, min(4000, int(1300.0 / max(est[0].elapsed_time(est[1]), 1e-3))))   # >= 1.3 s: a few clock samples per row
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(0) as clocks:
            e0.record()
            for _ in range(n):
                eng.replay_step(B, a.replay_size, it); it += 1
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        row = {"mode": mode, "batch": B, "ms_per_step": round(ms, 4), "steps_per_s": round(1000 / ms, 1),
               "samples_per_s": round(B * 1000 / ms), "tflops_algorithmic": round(FLOP * B / ms / 1e9, 2),
               "steps": n, "finite": bool(torch.isfinite(eng.params).all()), "clocks": clocks.summary()}
        print(json.dumps(row), flush=True)
        out.append(row)
        eng.close(); del eng
        torch.cuda.empty_cache()
