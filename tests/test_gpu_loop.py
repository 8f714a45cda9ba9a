"""Loop parity (BASELINE.json configs[0]): drop-in trainer + device replay ring + CUDA update engine vs the
unmodified reference loop on the same stand-in Pendulum, same seed (fixture tests/golden/loop_pendulum.npz from
tests/golden/make_golden_loop.py).  north_star's gate: first 100 losses and the returns within 1e-4 relative."""
import os
import random

import numpy as np
import pytest
import torch

from dsac_v2_b200 import synth
from loop.standin import Evaluator, Sampler, loop_kwargs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gemm", ["fp32", "bf16x3"])
def test_training_loop_tracks_reference(golden_dir, tmp_path, gemm):
    import dsac_v2
    from training.replay_buffer import ReplayBuffer
    from training.trainer import create_trainer
    z = np.load(os.path.join(golden_dir, "loop_pendulum.npz"))
    seed, iters = int(z["meta"][0]), int(z["meta"][1])
    keys = [str(k) for k in z["tb_keys"]]
    args = loop_kwargs(synth.reference_kwargs(synth.CONFIGS["pendulum"]), seed, str(tmp_path), max_iteration=iters,
                       dsact_noise="reference", dsact_gemm=gemm, dsact_tensorboard=False)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)      # reference utils/common_utils.py:140-156
    alg = dsac_v2.DSAC_V2(**args)
    sampler = Sampler(dsac_v2.ApproxContainer, **args)
    buffer = ReplayBuffer(**args)
    evaluator = Evaluator(dsac_v2.ApproxContainer, **args)
    rec = []
    inner = alg.local_update

    def local_update(data, it):
        tb = inner(data, it)
        rec.append([tb[k] for k in keys])
        return tb

    alg.local_update = local_update
    trainer = create_trainer(alg, sampler, buffer, evaluator, **args)
    assert buffer.size == 1000                                            # warm-up: 50 sampler calls of 20 steps
    for _ in range(iters):
        trainer.step()
        trainer.iteration += 1
    got, ref = np.array(rec), z["tb"]
    assert buffer.size == int(z["meta"][2]) and sampler.get_total_sample_number() == int(z["meta"][3])
    worst = np.max(np.abs(got[:100] - ref[:100]) / np.maximum(np.abs(ref[:100]), 1e-2))
    print(f"{gemm}: worst relative tb_info deviation over the first 100 iterations of the loop = {worst:.2e}")
    np.testing.assert_allclose(got[:100], ref[:100], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got[100:], ref[100:], rtol=5e-4, atol=1e-6)
    rets = np.array(evaluator.returns)
    np.testing.assert_allclose(rets[:, 0], z["returns"][:, 0])
    np.testing.assert_allclose(rets[:, 1], z["returns"][:, 1], rtol=1e-4)
