"""Host-side helpers shared by the drop-in modules (`dsac_v2`, `networks.mlp`,
`training.*`): kwargs plumbing, the action distributions the CPU sampler and
evaluator call, and the TensorBoard tag names.  None of this is on the update
path; the update itself runs in libdsact.so.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.append(_REPO)  # for the `dsac_v2_b200` import shim at the repo root

EPS = 1e-6  # reference utils/act_distribution_cls.py:3

# tag names of reference utils/tensorboard_setup.py:142-153 that the update path emits
try:  # a DSAC-v2 checkout on sys.path provides the full table
    from utils.tensorboard_setup import tb_tags as TB_TAGS  # type: ignore
except Exception:  # standalone
    TB_TAGS = {
        "loss_actor": "Loss/Actor loss-RL iter",
        "loss_critic": "Loss/Critic loss-RL iter",
        "alg_time": "Time/Algorithm time [ms]-RL iter",
        "sampler_time": "Time/Sampler time [ms]-RL iter",
        "TAR of RL iteration": "Evaluation/1. TAR-RL iter",
        "TAR of total time": "Evaluation/2. TAR-Total time [s]",
        "TAR of collected samples": "Evaluation/3. TAR-Collected samples",
        "TAR of replay samples": "Evaluation/4. TAR-Replay samples",
        "Buffer RAM of RL iteration": "RAM/RAM [MB]-RL iter",
    }


class _DiagGaussBase:
    """Shared pieces of the two diagonal-Gaussian action distributions
    (reference utils/act_distribution_cls.py:21-116).  `logits` = cat(mean, std)."""

    def __init__(self, logits: torch.Tensor):
        self.logits = logits
        self.mean, self.std = torch.chunk(logits, 2, dim=-1)
        self.act_high_lim = torch.tensor([1.0])
        self.act_low_lim = torch.tensor([-1.0])

    def _gauss_logp(self, x):
        z = (x - self.mean) / self.std
        return (-0.5 * z * z - self.std.log() - 0.5 * math.log(2 * math.pi)).sum(-1)

    def _draw(self, reparam: bool):
        noise = torch.randn_like(self.mean)
        x = self.mean + self.std * noise
        return x if reparam else x.detach()

    def entropy(self):
        return (0.5 + 0.5 * math.log(2 * math.pi) + self.std.log()).sum(-1)

    def kl_divergence(self, other):
        var_ratio = (self.std / other.std) ** 2
        t1 = ((self.mean - other.mean) / other.std) ** 2
        return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(-1)


class TanhGaussDistribution(_DiagGaussBase):
    """a = scale*tanh(u)+shift, u ~ N(mean, std) (reference :21-79)."""

    def _squash(self, u):
        scale = (self.act_high_lim - self.act_low_lim) / 2
        shift = (self.act_high_lim + self.act_low_lim) / 2
        logp = self._gauss_logp(u) - torch.log(1 + EPS - torch.tanh(u) ** 2).sum(-1) - torch.log(scale).sum(-1)
        return scale * torch.tanh(u) + shift, logp

    def sample(self):
        return self._squash(self._draw(False))

    def rsample(self):
        return self._squash(self._draw(True))

    def log_prob(self, action_limited):
        span = self.act_high_lim - self.act_low_lim
        u = torch.atanh((1 - EPS) * (2 * action_limited - (self.act_high_lim + self.act_low_lim)) / span)
        return self._gauss_logp(u) - torch.log(span * (1 + EPS - torch.tanh(u) ** 2)).sum(-1)

    def mode(self):
        scale = (self.act_high_lim - self.act_low_lim) / 2
        return scale * torch.tanh(self.mean) + (self.act_high_lim + self.act_low_lim) / 2


class GaussDistribution(_DiagGaussBase):
    """Unsquashed variant (reference :82-116); not used by the DSAC-T update engine."""

    def sample(self):
        x = self._draw(False)
        return x, self._gauss_logp(x)

    def rsample(self):
        x = self._draw(True)
        return x, self._gauss_logp(x)

    def log_prob(self, action):
        return self._gauss_logp(action)

    def mode(self):
        return torch.clamp(self.mean, self.act_low_lim, self.act_high_lim)


DISTRIBUTIONS = {"TanhGaussDistribution": TanhGaussDistribution, "GaussDistribution": GaussDistribution}


class ActionDistributionMixin:
    """`get_act_dist(logits)` of reference utils/act_distribution_cls.py:9-18."""

    def get_act_dist(self, logits):
        dist = self.action_distribution_cls(logits)
        if hasattr(self, "act_high_lim"):
            dist.act_high_lim = self.act_high_lim
            dist.act_low_lim = self.act_low_lim
        return dist


def net_kwargs(kind: str, kwargs: dict) -> dict:
    """Per-network constructor arguments out of the flat kwargs dict; same keys and
    defaults as reference utils/common_utils.py:48-89 (MLP branch)."""
    func_type = kwargs[kind + "_func_type"]
    if func_type not in ("MLP", "CNN"):
        raise NotImplementedError(f"{kind}_func_type={func_type!r}: MLP and CNN networks run on the B200 engine (not CNN_SHARED)")
    if kwargs.get("action_type", "continu") != "continu":
        raise NotImplementedError("DSAC don't support discrete action space!")
    dist = kwargs.get("policy_act_distribution", "TanhGaussDistribution")
    cls = dist if isinstance(dist, type) else DISTRIBUTIONS.get(dist)
    if cls is None:
        raise NotImplementedError(f"unknown action distribution {dist!r}")
    extra = dict(hidden_sizes=list(kwargs[kind + "_hidden_sizes"])) if func_type == "MLP" else dict(conv_type=kwargs[kind + "_conv_type"])
    return dict(
        apprfunc=func_type,
        name=kwargs[kind + "_func_name"],
        obs_dim=kwargs["obsv_dim"],
        act_dim=kwargs["action_dim"],
        **extra,
        hidden_activation=kwargs[kind + "_hidden_activation"],
        output_activation=kwargs[kind + "_output_activation"],
        min_log_std=kwargs.get(kind + "_min_log_std", -20.0),
        max_log_std=kwargs.get(kind + "_max_log_std", 2.0),
        std_type=kwargs.get(kind + "_std_type", "mlp_shared"),
        act_high_lim=np.array(kwargs["action_high_limit"], dtype=np.float32),
        act_low_lim=np.array(kwargs["action_low_limit"], dtype=np.float32),
        action_distribution_cls=cls,
    )
