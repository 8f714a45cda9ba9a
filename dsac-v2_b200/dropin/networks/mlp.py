"""`networks.mlp` of the drop-in: the two MLP approximators DSAC-T uses, with the
reference's class names, constructor kwargs and parameter names
(reference networks/mlp.py:28-127), so `state_dict()` keeps the shipped
checkpoint schema (`policy.policy.0.weight`, `q1.q.0.weight`, ...).

These modules are containers + the plain-torch forward used by the CPU sampler,
the evaluator and checkpoint tools.  During training their parameters are views
into the engine's flat device buffers (see `dsac_v2.ApproxContainer`), and the
update path never calls these forwards: it runs in libdsact.so.
"""
__all__ = ["StochaPolicy", "ActionValueDistri"]

import torch
import torch.nn as nn

from dsact_host import ActionDistributionMixin

_ACTIVATIONS = {"relu": nn.ReLU, "elu": nn.ELU, "gelu": nn.GELU, "selu": nn.SELU, "sigmoid": nn.Sigmoid,
                "tanh": nn.Tanh, "linear": nn.Identity}


def _activation(name):
    if isinstance(name, str) and name in _ACTIVATIONS:
        return _ACTIVATIONS[name]
    print("input activation name:" + str(name))
    raise RuntimeError


def build_mlp(sizes, hidden_activation, output_activation):
    """Linear/activation pairs; even indices are the Linear layers (checkpoint key numbering)."""
    hidden, out = _activation(hidden_activation), _activation(output_activation)
    mods = []
    last = len(sizes) - 2
    for j, (n_in, n_out) in enumerate(zip(sizes[:-1], sizes[1:])):
        mods.append(nn.Linear(n_in, n_out))
        mods.append(out() if j == last else hidden())
    return nn.Sequential(*mods)


class StochaPolicy(nn.Module, ActionDistributionMixin):
    """obs -> cat(mean, std) of the action distribution; std = exp(clamp(log_std))."""

    def __init__(self, **kwargs):
        super().__init__()
        self.std_type = kwargs["std_type"]
        obs_dim, act_dim = kwargs["obs_dim"], kwargs["act_dim"]
        net = lambda out: build_mlp([obs_dim, *kwargs["hidden_sizes"], out], kwargs["hidden_activation"], kwargs["output_activation"])
        if self.std_type == "mlp_shared":          # one MLP, 2 * act_dim outputs (reference :56-62)
            self.policy = net(2 * act_dim)
        elif self.std_type == "mlp_separated":     # two MLPs (reference :43-54); creation order = the reference's RNG consumption
            self.mean = net(act_dim)
            self.log_std = net(act_dim)
        elif self.std_type == "parameter":         # mean MLP + learnable row (reference :64-71)
            self.mean = net(act_dim)
            self.log_std = nn.Parameter(-0.5 * torch.ones(1, act_dim))
        else:
            raise NotImplementedError(f"policy std_type={self.std_type!r}")
        self.min_log_std = kwargs["min_log_std"]
        self.max_log_std = kwargs["max_log_std"]
        self.register_buffer("act_high_lim", torch.from_numpy(kwargs["act_high_lim"]))
        self.register_buffer("act_low_lim", torch.from_numpy(kwargs["act_low_lim"]))
        self.action_distribution_cls = kwargs["action_distribution_cls"]

    def forward(self, obs):
        if self.std_type == "mlp_shared":
            mean, log_std = self.policy(obs).chunk(2, dim=-1)
        elif self.std_type == "mlp_separated":
            mean, log_std = self.mean(obs), self.log_std(obs)
        else:
            mean = self.mean(obs)
            log_std = self.log_std + torch.zeros_like(mean)
        std = log_std.clamp(self.min_log_std, self.max_log_std).exp()
        return torch.cat((mean, std), dim=-1)


class ActionValueDistri(nn.Module):
    """(obs, act) -> cat(mean, softplus(std)) of the return distribution."""

    def __init__(self, **kwargs):
        super().__init__()
        self.q = build_mlp([kwargs["obs_dim"] + kwargs["act_dim"], *kwargs["hidden_sizes"], 2],
                           kwargs["hidden_activation"], kwargs["output_activation"])

    def forward(self, obs, act):
        mean, raw_std = self.q(torch.cat([obs, act], dim=-1)).chunk(2, dim=-1)
        return torch.cat((mean, nn.functional.softplus(raw_std)), dim=-1)
