"""Deterministic synthetic inputs for the DSAC-T update path (numpy only).

Everything the parity tests, the golden-vector generator, ``smoke()`` and
``bench.py`` feed to the update step comes from here, so that the reference
run (``tests/golden/make_golden.py``, executed once in the build container)
and the CUDA run (on the GPU box, where the reference does not exist) see
bit-identical weights, minibatches and noise.  Streams are numpy PCG64
(`numpy.random.default_rng`), whose output is stable across platforms.

Shapes follow SURVEY.md §8(d): obs/obs2 ~ N(0,1) [B,O], act ~ U(lo,hi) [B,A],
rew ~ N(0,1) [B], done ~ Bernoulli(0.01) [B]; noise = the eight normal draws
one `DSAC_V2.local_update` consumes (reference dsac_v2.py:160,228,212 — see
SURVEY.md Appendix B).
"""
from __future__ import annotations

import numpy as np

# Named problem shapes of BASELINE.json `configs` (P, H, C) plus a tiny one
# whose complete state fits in a golden fixture.
CONFIGS = {
    "tiny": dict(obs_dim=5, act_dim=2, hidden=(32, 32), act_lim=1.0),
    "pendulum": dict(obs_dim=3, act_dim=1, hidden=(256, 256, 256), act_lim=2.0),
    "halfcheetah": dict(obs_dim=17, act_dim=6, hidden=(256, 256), act_lim=1.0),
    "humanoid": dict(obs_dim=376, act_dim=17, hidden=(256, 256, 256), act_lim=0.4),
    # ragged: nothing is a multiple of the GEMM tile or of the vector width
    "ragged": dict(obs_dim=11, act_dim=3, hidden=(40, 24, 72), act_lim=1.5),
}

# BASELINE.json config 5 (gym_carracingraw): conv encoder `type_2` + separate mean / log_std
# heads (csrc/cnn_engine.cuh; oracle/dsact_oracle.py:OracleDSACTCNN, tests/golden/cnn_carracing_b4.npz).
CNN_CONFIGS = {
    "carracing": dict(obs_dim=(3, 96, 96), act_dim=3, act_lim=1.0, conv_type="type_2"),
    # the reference's other encoder (networks/cnn.py:173-186: 8x8/4, 4x4/2, 3x3/1, heads 512-256) on a smaller image
    "small_t1": dict(obs_dim=(2, 44, 44), act_dim=2, act_lim=1.0, conv_type="type_1"),
    "odd": dict(obs_dim=(3, 13, 11), act_dim=2, act_lim=1.0, conv_type="test_odd"),
}
# reference networks/cnn.py:201-216 (type_2) and :163-170 (type_1): kernel sizes, channels, strides, head widths
CONV_TYPES = {
    "type_1": dict(kernels=(8, 4, 3), channels=(32, 64, 64), strides=(4, 2, 1), heads=(512, 256)),
    "type_2": dict(kernels=(4, 3, 3, 3, 3, 3), channels=(8, 16, 32, 64, 128, 256), strides=(2, 2, 2, 2, 1, 1), heads=(256, 256, 256)),
    # not a reference type: channel counts that are no multiples of 4 / 8 and 2x2 / 1x1 windows, for the one-channel-per-thread
    # kernels of conv.cuh (tests only; checked against the oracle)
    "test_odd": dict(kernels=(3, 2, 1), channels=(6, 10, 12), strides=(2, 1, 1), heads=(24,)),
}

HYPER = dict(
    gamma=0.99,
    tau=0.005,
    delay_update=2,
    auto_alpha=True,
    alpha=0.2,
    value_learning_rate=1e-4,
    policy_learning_rate=1e-4,
    alpha_learning_rate=3e-4,
    policy_min_log_std=-20.0,
    policy_max_log_std=0.5,
)


def _rng(*key) -> np.random.Generator:
    return np.random.default_rng([int(k) for k in key])


def net_shapes(obs_dim, act_dim, hidden):
    """(q_sizes, pi_sizes) layer-size lists, reference networks/mlp.py:58,116."""
    q = [obs_dim + act_dim] + list(hidden) + [2]
    pi = [obs_dim] + list(hidden) + [2 * act_dim]
    return q, pi


def make_weights(cfg: dict, seed: int = 0) -> dict:
    """state_dict-shaped fp32 arrays for q1, q2, policy (targets = copies).

    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default init; keys use
    the reference's schema (`q1.q.{0,2,..}.weight`, `policy.policy.{0,2,..}.bias`).
    """
    q_sizes, pi_sizes = net_shapes(cfg["obs_dim"], cfg["act_dim"], cfg["hidden"])
    out = {}
    for n, (name, inner, sizes) in enumerate(
        (("q1", "q", q_sizes), ("q2", "q", q_sizes), ("policy", "policy", pi_sizes))
    ):
        g = _rng(seed, 11, n)
        for j in range(len(sizes) - 1):
            bound = 1.0 / np.sqrt(sizes[j])
            out[f"{name}.{inner}.{2 * j}.weight"] = g.uniform(
                -bound, bound, size=(sizes[j + 1], sizes[j])
            ).astype(np.float32)
            out[f"{name}.{inner}.{2 * j}.bias"] = g.uniform(
                -bound, bound, size=(sizes[j + 1],)
            ).astype(np.float32)
    for src, dst in (("q1", "q1_target"), ("q2", "q2_target"), ("policy", "policy_target")):
        for k in [k for k in out if k.startswith(src + ".")]:
            out[dst + k[len(src):]] = out[k].copy()
    return out


def make_batch(cfg: dict, batch: int, step: int, seed: int = 123) -> dict:
    g = _rng(seed, 22, step)
    O, A, lim = cfg["obs_dim"], cfg["act_dim"], cfg["act_lim"]
    return {
        "obs": g.standard_normal((batch, O)).astype(np.float32),
        "act": g.uniform(-lim, lim, size=(batch, A)).astype(np.float32),
        "rew": g.standard_normal(batch).astype(np.float32),
        "obs2": g.standard_normal((batch, O)).astype(np.float32),
        "done": (g.random(batch) < 0.01).astype(np.float32),
        "logp": np.zeros(batch, dtype=np.float32),
    }


def make_noise(cfg: dict, batch: int, step: int, seed: int = 7) -> list:
    """The 8 standard-normal draws of one update, in the reference's order:
    eps1 [B,A], eps2 [B,A], z1..z6 [B] (z3,z4 are the two that matter)."""
    g = _rng(seed, 33, step)
    A = cfg["act_dim"]
    out = [g.standard_normal((batch, A)).astype(np.float32) for _ in range(2)]
    out += [g.standard_normal(batch).astype(np.float32) for _ in range(6)]
    return out


def reference_kwargs(cfg: dict, **over) -> dict:
    """The kwargs dict the reference threads through DSAC_V2 / ApproxContainer
    (what example_train/main.py + utils/init_args.py would have produced)."""
    lim = np.full(cfg["act_dim"], cfg["act_lim"], dtype=np.float32)
    kw = dict(
        algorithm="DSAC_V2",
        obsv_dim=cfg["obs_dim"],
        action_dim=cfg["act_dim"],
        action_type="continu",
        action_high_limit=lim,
        action_low_limit=-lim,
        value_func_name="ActionValueDistri",
        value_func_type="MLP",
        value_hidden_sizes=list(cfg["hidden"]),
        value_hidden_activation="gelu",
        value_output_activation="linear",
        policy_func_name="StochaPolicy",
        policy_func_type="MLP",
        policy_act_distribution="TanhGaussDistribution",
        policy_hidden_sizes=list(cfg["hidden"]),
        policy_hidden_activation="gelu",
        policy_output_activation="linear",
        cnn_shared=False,
    )
    kw.update(HYPER)
    kw.update(over)
    return kw


# ---- CNN variant (config 5) -----------------------------------------------------------------------------------------
def conv_feature_dim(cfg: dict) -> int:
    c, h, w = cfg["obs_dim"]
    t = CONV_TYPES[cfg["conv_type"]]
    for k, s_ in zip(t["kernels"], t["strides"]):
        h, w = (h - k) // s_ + 1, (w - k) // s_ + 1
    return t["channels"][-1] * h * w


def make_cnn_weights(cfg: dict, seed: int = 0) -> dict:
    """state_dict-shaped fp32 arrays in the schema of the reference's CNN `StochaPolicy` / `ActionValueDistri`
    (networks/cnn.py:151-240, 383-461): `{net}.conv.{0,2,..}.{weight,bias}`, `{net}.mean.{0,2,..}`, `{net}.log_std.{0,2,..}`;
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like the torch defaults; targets are copies."""
    t = CONV_TYPES[cfg["conv_type"]]
    feat, A = conv_feature_dim(cfg), cfg["act_dim"]
    out = {}
    for n, (net, extra, width) in enumerate((("q1", A, 1), ("q2", A, 1), ("policy", 0, A))):
        g = _rng(seed, 44, n)
        cin = cfg["obs_dim"][0]
        for j, (k, cout) in enumerate(zip(t["kernels"], t["channels"])):
            bound = 1.0 / np.sqrt(cin * k * k)
            out[f"{net}.conv.{2 * j}.weight"] = g.uniform(-bound, bound, (cout, cin, k, k)).astype(np.float32)
            out[f"{net}.conv.{2 * j}.bias"] = g.uniform(-bound, bound, (cout,)).astype(np.float32)
            cin = cout
        sizes = [feat + extra] + list(t["heads"]) + [width]
        for head in ("mean", "log_std"):
            for j in range(len(sizes) - 1):
                bound = 1.0 / np.sqrt(sizes[j])
                out[f"{net}.{head}.{2 * j}.weight"] = g.uniform(-bound, bound, (sizes[j + 1], sizes[j])).astype(np.float32)
                out[f"{net}.{head}.{2 * j}.bias"] = g.uniform(-bound, bound, (sizes[j + 1],)).astype(np.float32)
    for src, dst in (("q1", "q1_target"), ("q2", "q2_target"), ("policy", "policy_target")):
        for k in [k for k in out if k.startswith(src + ".")]:
            out[dst + k[len(src):]] = out[k].copy()
    return out


def make_cnn_batch(cfg: dict, batch: int, step: int, seed: int = 123) -> dict:
    """Image minibatch: obs/obs2 ~ U(0,1) [B,C,H,W] (pixel-like), the rest as `make_batch`."""
    g = _rng(seed, 55, step)
    A, lim = cfg["act_dim"], cfg["act_lim"]
    shape = (batch,) + tuple(cfg["obs_dim"])
    return {
        "obs": g.random(shape, dtype=np.float32),
        "act": g.uniform(-lim, lim, size=(batch, A)).astype(np.float32),
        "rew": g.standard_normal(batch).astype(np.float32),
        "obs2": g.random(shape, dtype=np.float32),
        "done": (g.random(batch) < 0.01).astype(np.float32),
    }


def cnn_reference_kwargs(cfg: dict, **over) -> dict:
    """kwargs of example_train/dsacv2_cnn_carracing_offasync.py:53-79 for `DSAC_V2` / `ApproxContainer`."""
    kw = reference_kwargs(dict(obs_dim=0, act_dim=cfg["act_dim"], act_lim=cfg["act_lim"], hidden=()))
    for k in ("value_hidden_sizes", "policy_hidden_sizes"):
        kw.pop(k)
    kw.update(obsv_dim=tuple(cfg["obs_dim"]), value_func_type="CNN", policy_func_type="CNN",
              value_conv_type=cfg["conv_type"], policy_conv_type=cfg["conv_type"])
    kw.update(over)
    return kw


# ---- other policy std types (reference networks/mlp.py:42-72) ---------------------------------
def make_weights_std(cfg: dict, std_type: str, seed: int = 0) -> dict:
    """`make_weights` with the policy in the schema of `std_type`:
    "mlp_separated": `policy.mean.{0,2,..}` and `policy.log_std.{0,2,..}` (two MLPs ending in act_dim outputs),
    "parameter": `policy.mean.{0,2,..}` and the learnable row `policy.log_std` [1, act_dim] (= -0.5)."""
    out = {k: v for k, v in make_weights(cfg, seed).items() if not k.startswith("policy")}
    sizes = [cfg["obs_dim"]] + list(cfg["hidden"]) + [cfg["act_dim"]]
    g = _rng(seed, 66, {"mlp_separated": 0, "parameter": 1}[std_type])
    heads = ("mean", "log_std") if std_type == "mlp_separated" else ("mean",)
    for head in heads:
        for j in range(len(sizes) - 1):
            bound = 1.0 / np.sqrt(sizes[j])
            out[f"policy.{head}.{2 * j}.weight"] = g.uniform(-bound, bound, (sizes[j + 1], sizes[j])).astype(np.float32)
            out[f"policy.{head}.{2 * j}.bias"] = g.uniform(-bound, bound, (sizes[j + 1],)).astype(np.float32)
    if std_type == "parameter":
        out["policy.log_std"] = np.full((1, cfg["act_dim"]), -0.5, dtype=np.float32)
    for k in [k for k in out if k.startswith("policy.")]:
        out["policy_target" + k[len("policy"):]] = out[k].copy()
    return out


# ---- DSAC_V1 (reference dsac_v1.py) --------------------------------------------------------------
def make_weights_v1(cfg: dict, seed: int = 0) -> dict:
    """`make_weights` in the schema of `dsac_v1.ApproxContainer` (:17-52): one critic `q.q.*`, `policy.policy.*`, targets."""
    w = make_weights(cfg, seed)
    out = {}
    for k, v in w.items():
        if k.startswith("q1"):
            out["q" + k[2:]] = v.copy()
        elif k.startswith("policy"):
            out[k] = v.copy()
    return out
