"""Iterations per second through the drop-in `OffSerialTrainer.step()` — the whole loop glue, not only `local_update`:
CPU env stepping (stand-in Pendulum of tests/loop), replay `add_batch` to the device ring, device gather, update, policy
mirror.  Three modes: the reference's serial semantics (sample 20 env steps with the current policy every iteration),
the serial loop with a relaxed mirror interval, and the asynchronous sampler thread (north_star: "the env loop stays on
CPU and feeds the buffer asynchronously").

    python tools/trainer_rate.py [--iters 2000] [--batch 256]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dsac-v2_b200", "dropin"), ROOT, os.path.join(ROOT, "tests")]

import torch  # noqa: E402

torch.set_num_threads(4)   # utils/init_args.py:14 of the reference runs torch's CPU side with 4 threads; batch-1 policy
                           # forwards on every core of a 128-thread host cost 15 ms each instead of 0.1 ms

from dsac_v2_b200 import synth  # noqa: E402
from loop.standin import Evaluator, Sampler, loop_kwargs  # noqa: E402


def run(mode, iters, batch, min_seconds=3.0):
    import dsac_v2
    from training.replay_buffer import ReplayBuffer
    from training.trainer import create_trainer
    over = dict(max_iteration=10 ** 9, eval_interval=10 ** 9, log_save_interval=10 ** 9, apprfunc_save_interval=10 ** 9,
                replay_batch_size=batch, buffer_warm_size=max(1000, batch), dsact_tensorboard=False)
    if mode == "serial_mirror50":
        over["policy_mirror_interval"] = 50
    if mode == "async":
        over["dsact_async_sampler"] = True
    with tempfile.TemporaryDirectory() as tmp:
        args = loop_kwargs(synth.reference_kwargs(synth.CONFIGS["pendulum"]), 1, tmp, **over)
        torch.manual_seed(1)
        alg = dsac_v2.DSAC_V2(**args)
        sampler, buffer = Sampler(dsac_v2.ApproxContainer, **args), ReplayBuffer(**args)
        trainer = create_trainer(alg, sampler, buffer, Evaluator(dsac_v2.ApproxContainer, **args), **args)
        trainer.iteration = 1                      # iteration 0 evaluates / logs / saves
        for _ in range(50):
            trainer.step(); trainer.iteration += 1
        torch.cuda.synchronize()
        n0, t0 = sampler.get_total_sample_number(), time.perf_counter()
        done = 0
        while done < iters or time.perf_counter() - t0 < min_seconds:
            trainer.step(); trainer.iteration += 1
            done += 1
        iters = done
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        env_steps = sampler.get_total_sample_number() - n0
        trainer.close()
    return {"mode": mode, "replay_batch_size": batch, "iterations": iters, "iters_per_s": round(iters / dt, 1),
            "ms_per_iter": round(1e3 * dt / iters, 4), "env_steps_per_s": round(env_steps / dt, 1),
            "env_steps_per_iter": round(env_steps / iters, 2)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    for mode in ("serial", "serial_mirror50", "async"):
        print(json.dumps(run(mode, a.iters if mode == "async" else max(200, a.iters // 10), a.batch)), flush=True)
