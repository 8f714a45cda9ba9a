"""DSAC_V1 (reference dsac_v1.py) through the C ABI: the head-wise fp32 engine with `algo = 1` (one
critic, flat layout [q | policy | log_alpha], fixed TD bound) against the goldens produced by the unmodified reference
(tests/golden/v1_*.npz: bounded loss, a tight bound that clips, the Gaussian NLL), against the pinned oracle with gradients,
and through the drop-in `dsac_v1.DSAC_V1`."""
import ast
import os

import numpy as np
import pytest
import torch

from dsac_v2_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-4
# columns of the engine's 16 statistics that carry DSAC_V1's tb_info (dsac_v1.py:172-181), in V1_TB_KEYS order
V1_COLS = [0, 2, 6, 8, 9, 10, 11]


def make_engine(cfg, batch, over):
    from dsac_v2_b200.engine_cnn import CnnEngine, make_heads_config
    h = dict(synth.HYPER)
    h.update(over)
    c = make_heads_config(cfg["obs_dim"], cfg["act_dim"], cfg["hidden"], "mlp_shared", max_batch=batch, algo="DSAC_V1",
                          bound=h.get("bound", True), td_bound=h.get("TD_bound", 20), gamma=h["gamma"], tau=h["tau"],
                          delay_update=h["delay_update"], auto_alpha=h["auto_alpha"], alpha=h["alpha"], lr_q=h["value_learning_rate"],
                          lr_pi=h["policy_learning_rate"], lr_alpha=h["alpha_learning_rate"], min_log_std=h["policy_min_log_std"],
                          max_log_std=h["policy_max_log_std"])
    lim = torch.full((cfg["act_dim"],), cfg["act_lim"])
    eng = CnnEngine(c, torch.device("cuda", 0), lim, -lim)
    eng.load_weights(synth.make_weights_v1(cfg))
    return eng


def feed(cfg, batch, it):
    b = {k: torch.from_numpy(v).cuda() for k, v in synth.make_batch(cfg, batch, it).items()}
    n = synth.make_noise(cfg, batch, it)
    return b, tuple(torch.from_numpy(n[i]).cuda() for i in (0, 1, 3, 3))   # eps1, eps2, the target critic's z (twice)


def stats_v1(eng):
    from dsac_v2_b200.engine import STAT_KEYS
    s = eng.read_stats()
    v = [s[k] for k in STAT_KEYS]
    return np.array([v[i] for i in V1_COLS])


@pytest.mark.parametrize("name", ["v1_tiny_b16", "v1_ragged_tight", "v1_tiny_nll"])
def test_v1_update_matches_reference_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg_name, batch, steps, over = z["meta"]
    cfg, batch, steps, over = synth.CONFIGS[str(cfg_name)], int(batch), int(steps), dict(ast.literal_eval(str(over)))
    eng = make_engine(cfg, batch, over)
    names = [str(n) for n in z["param_names"]]
    for it in range(steps):
        b, n = feed(cfg, batch, it)
        eng.step(b, it, n)
        np.testing.assert_allclose(stats_v1(eng), z["tb"][it], rtol=RTOL, atol=1e-6, err_msg=f"{name} tb_info at step {it}")
        if f"pdigest_{it + 1}" in z:
            w = eng.export_weights()
            for row, k in zip(z[f"pdigest_{it + 1}"], names):
                d = w[k].double().reshape(-1)
                np.testing.assert_allclose(d.abs().sum().item(), row[1], rtol=RTOL, err_msg=f"{name} {k} step {it + 1}")
                np.testing.assert_allclose(d[:8].numpy(), row[3:3 + min(8, d.numel())], rtol=RTOL, atol=1e-7, err_msg=f"{name} {k} step {it + 1}")
        if f"state_{it + 1}/{names[0]}" in z:
            w = eng.export_weights()
            for k in names:
                ref = z[f"state_{it + 1}/{k}"]
                np.testing.assert_allclose(w[k].numpy(), ref, rtol=RTOL, atol=1e-6 * max(1e-3, np.abs(ref).max()), err_msg=f"{name} {k} after step {it + 1}")
    eng.close()


@pytest.mark.parametrize("cfg_name,batch,over", [("ragged", 50, {}), ("tiny", 1, {"TD_bound": 1.0}), ("pendulum", 300, {"bound": False})])
def test_v1_update_matches_oracle(cfg_name, batch, over):
    from oracle.dsact_oracle import V1_TB_KEYS, v1_from_config
    cfg = synth.CONFIGS[cfg_name]
    eng = make_engine(cfg, batch, over)
    hyper = dict(synth.HYPER)
    hyper.update(over)
    orc = v1_from_config(cfg, synth.make_weights_v1(cfg), **hyper)
    for it in range(4):
        ref = orc.update(synth.make_batch(cfg, batch, it), synth.make_noise(cfg, batch, it), it)
        b, n = feed(cfg, batch, it)
        eng.step(b, it, n)
        np.testing.assert_allclose(stats_v1(eng), [ref[k] for k in V1_TB_KEYS], rtol=RTOL, atol=1e-6, err_msg=f"step {it}")
    g, gref = eng.export_weights(grads=True), orc.grad_dict()
    for k, v in gref.items():
        np.testing.assert_allclose(g[k].numpy(), v.numpy(), rtol=1e-3, atol=2e-6 * float(v.abs().max()) + 1e-12, err_msg=f"grad {k}")
    w, sd = eng.export_weights(), orc.state_dict()
    for k, v in sd.items():
        np.testing.assert_allclose(w[k].numpy(), v.numpy(), rtol=RTOL, atol=1e-5, err_msg=k)
    eng.close()


def test_v1_dropin_local_update(golden_dir):
    """`dsac_v1.DSAC_V1(**kwargs)`: the reference's container (q, q_target, policy, policy_target, log_alpha) with its
    state_dict keys and parameter order, `local_update` on the GPU returning the reference's tb_info keys."""
    import dsac_v1
    from oracle.dsact_oracle import V1_TB_KEYS
    z = np.load(os.path.join(golden_dir, "v1_tiny_b16.npz"))
    cfg, B = synth.CONFIGS["tiny"], int(z["meta"][1])
    kw = synth.reference_kwargs(cfg, algorithm="DSAC_V1", replay_batch_size=B)
    alg = dsac_v1.DSAC_V1(**kw)
    sd = alg.networks.state_dict()
    ref_w = synth.make_weights_v1(cfg)
    assert {k for k in sd if not k.endswith("_lim")} == set(ref_w) | {"log_alpha"}
    assert [k for k, _ in alg.networks.named_parameters()] == [str(n) for n in z["param_names"]]
    for k, v in ref_w.items():
        sd[k] = torch.from_numpy(v)
    alg.networks.load_state_dict(sd)
    alg.networks.cuda()
    eng = alg.networks.engine(B)
    for it in range(3):
        b, n = feed(cfg, B, it)
        eng.step(b, it, n)
        np.testing.assert_allclose(stats_v1(eng), z["tb"][it], rtol=RTOL, atol=1e-6)
    assert not torch.equal(alg.networks.state_dict()["q.q.0.weight"].cpu(), torch.from_numpy(ref_w["q.q.0.weight"]))
    tb = alg.local_update({k: torch.from_numpy(v) for k, v in synth.make_batch(cfg, B, 9).items()}, 3)   # host minibatch
    assert set(V1_TB_KEYS) <= set(tb) and all(np.isfinite(tb[k]) for k in V1_TB_KEYS)
    with pytest.raises(NotImplementedError):
        alg.get_remote_update_info({}, 0)


@pytest.mark.parametrize("variant", ["DSAC_V1", "mlp_separated", "parameter"])
def test_trainer_loop_on_the_head_wise_engine(tmp_path, variant):
    """The drop-in `OffSerialTrainer` (device replay ring, policy mirror for the CPU sampler) around the head-wise engine:
    DSAC_V1 and DSAC-T with the policy's other std types."""
    import dsac_v1
    import dsac_v2
    from training.replay_buffer import ReplayBuffer
    from training.trainer import create_trainer
    cfg = synth.CONFIGS["tiny"]
    mod, over = (dsac_v1, {"algorithm": "DSAC_V1"}) if variant == "DSAC_V1" else (dsac_v2, {"policy_std_type": variant})
    kw = synth.reference_kwargs(cfg, replay_batch_size=32, **over)
    kw = dict(kw, buffer_max_size=1000, additional_info={}, buffer_name="replay_buffer", buffer_warm_size=100, max_iteration=12,
              log_save_interval=5, apprfunc_save_interval=10, eval_interval=6, save_folder=str(tmp_path), ini_network_dir=None,
              use_gpu=True, dsact_tensorboard=False)
    alg = (mod.DSAC_V1 if variant == "DSAC_V1" else mod.DSAC_V2)(**kw)

    class Sampler:
        def __init__(self):
            self.networks = mod.ApproxContainer(**kw)
            self.n, self.g = 0, np.random.default_rng(0)
            self.obs = self.g.standard_normal(cfg["obs_dim"]).astype(np.float32)

        def sample(self):
            out = []
            for _ in range(20):
                logits = self.networks.policy(torch.from_numpy(self.obs[None]))
                act, logp = self.networks.create_action_distributions(logits).sample()
                nxt = (0.9 * self.obs + 0.1 * self.g.standard_normal(cfg["obs_dim"])).astype(np.float32)
                out.append((self.obs.copy(), {}, act.detach()[0].numpy(), float(-np.abs(nxt).mean()), nxt.copy(), False, logp.detach()[0].numpy(), {}))
                self.obs = nxt
            self.n += 20
            return out, {}

        def get_total_sample_number(self):
            return self.n

    class Evaluator:
        networks, calls = None, 0

        def run_evaluation(self, it):
            self.calls += 1
            return 0.0

    sampler, evaluator = Sampler(), Evaluator()
    trainer = create_trainer(alg, sampler, ReplayBuffer(**kw), evaluator, **kw)
    first = next(iter(alg.networks.policy.parameters())).detach().clone()
    trainer.train()
    assert trainer.iteration == 12 and evaluator.calls == 2
    now = next(iter(alg.networks.policy.parameters())).detach()
    assert not torch.equal(first, now)
    trainer.refresh_policy_mirror()   # the CPU mirror the sampler acts with tracks the trained GPU policy
    torch.testing.assert_close(next(iter(sampler.networks.policy.parameters())).detach(), now.cpu(), rtol=0, atol=0)
    assert all(np.isfinite(float(v)) for v in trainer.last_tb.values())
