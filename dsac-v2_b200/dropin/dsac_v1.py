"""`dsac_v1` of the drop-in: `ApproxContainer` and `DSAC_V1` with the reference's names, kwargs and `tb_info` keys
(reference dsac_v1.py:17-52, 56-273) — the older algorithm (one distributional critic, fixed TD bound; selectable with
`--algorithm DSAC_V1`), backed by the head-wise fp32 engine of libdsact.so (`dsact_cnn_*` with `algo = 1`).

* `ApproxContainer`: `q`, `q_target`, `policy`, `policy_target` (the same `networks.mlp` / `networks.cnn` classes as
  DSAC-T) + `log_alpha`; on a CUDA device the parameters are views into the engine's flat buffers [q | policy | log_alpha].
* `DSAC_V1.local_update(data, iteration) -> tb_info` runs the whole update in the CUDA library; no CPU fallback.
  `get_remote_update_info` / `remote_update` (the gradient-message seam of the reference's asynchronous trainers) are not
  part of this engine and raise.

Extra kwargs: `dsact_noise` = "device" (default) | "reference" (draw eps1, eps2 and the three z's of one update from torch's
CPU generator in the reference's order), `dsact_max_batch`, `seed`.
"""
__all__ = ["ApproxContainer", "DSAC_V1"]

import time
from copy import deepcopy
from typing import Dict

import torch
import torch.nn as nn

import networks.cnn as _cnn
import networks.mlp as _mlp
from dsact_host import TB_TAGS as tb_tags
from dsact_host import net_kwargs

from dsac_v2_b200 import _lib
from dsac_v2_b200.engine_cnn import CnnEngine, make_cnn_config, make_heads_config

# where the engine's 16-slot statistics carry DSAC_V1's tb_info (dsac_v1.py:172-181)
_V1_KEYS = (("DSAC/critic_avg_q-RL iter", 0), ("DSAC/critic_avg_std-RL iter", 2), (tb_tags["loss_actor"], 6),
            ("DSAC/policy_mean-RL iter", 8), ("DSAC/policy_std-RL iter", 9), ("DSAC/entropy-RL iter", 10),
            ("DSAC/alpha-RL iter", 11))


class ApproxContainer(nn.Module):
    """One critic, one policy, their targets and log_alpha (reference dsac_v1.py:17-52)."""

    def __init__(self, **kwargs):
        super().__init__()
        q_args, pi_args = net_kwargs("value", kwargs), net_kwargs("policy", kwargs)
        if q_args["apprfunc"] != pi_args["apprfunc"]:
            raise NotImplementedError("value and policy approximators must be of the same type (both MLP or both CNN)")
        cnn = q_args["apprfunc"] == "CNN"
        if cnn:   # the engine's DSAC_V1 step is wired for encoders too, but only the MLP configuration is pinned to the reference
            raise NotImplementedError("DSAC_V1 on the B200 engine: MLP approximators (the CNN configuration is not validated)")
        if pi_args["std_type"] != "mlp_shared":   # same reason: wired in the engine (pi_std 0 / 1 with algo = 1), no reference golden
            raise NotImplementedError("DSAC_V1 on the B200 engine: policy std_type 'mlp_shared' (the reference's default for DSAC_V1)")
        mod = _cnn if cnn else _mlp
        q_cls, pi_cls = getattr(mod, q_args["name"], None), getattr(mod, pi_args["name"], None)
        if q_cls is None or pi_cls is None:
            raise NotImplementedError("This apprfunc is not properly defined")
        self.q = q_cls(**q_args)                      # construction order = the reference's RNG consumption (:28-34)
        self.q_target = deepcopy(self.q)
        self.policy = pi_cls(**pi_args)
        self.policy_target = deepcopy(self.policy)
        for net in (self.policy_target, self.q_target):
            for p in net.parameters():
                p.requires_grad = False
        self.log_alpha = nn.Parameter(torch.tensor(1, dtype=torch.float32))
        if q_args["output_activation"] != "linear" or pi_args["output_activation"] != "linear":
            raise NotImplementedError("the B200 engine implements linear output activations")
        if pi_args["action_distribution_cls"].__name__ not in _lib.ACT_DISTS:
            raise NotImplementedError("the B200 engine implements TanhGaussDistribution and GaussDistribution")
        if q_args["hidden_activation"] != pi_args["hidden_activation"]:
            raise NotImplementedError("the head-wise engine takes one hidden activation for critic and policy")
        common = dict(gamma=kwargs.get("gamma", 0.99), tau=kwargs.get("tau", 0.005), delay_update=kwargs.get("delay_update", 2),
                      auto_alpha=kwargs.get("auto_alpha", True), alpha=kwargs.get("alpha", 0.2), lr_q=kwargs["value_learning_rate"],
                      lr_pi=kwargs["policy_learning_rate"], lr_alpha=kwargs["alpha_learning_rate"],
                      min_log_std=pi_args["min_log_std"], max_log_std=pi_args["max_log_std"],
                      act_dist=pi_args["action_distribution_cls"].__name__, act_hidden=q_args["hidden_activation"],
                      algo="DSAC_V1", bound=kwargs.get("bound", True), td_bound=kwargs.get("TD_bound", 20))
        if cnn:
            if q_args["conv_type"] != pi_args["conv_type"]:
                raise NotImplementedError("the CNN engine takes one conv_type for critic and policy")
            t = _cnn.CONV_TYPES[q_args["conv_type"]]
            self._make = make_cnn_config
            self._cfg_args = dict(obs_shape=tuple(q_args["obs_dim"]), act_dim=q_args["act_dim"], kernels=t["kernels"],
                                  channels=t["channels"], strides=t["strides"], hidden=t["heads"], **common)
        else:
            if q_args["hidden_sizes"] != pi_args["hidden_sizes"]:
                raise NotImplementedError("the head-wise engine takes one hidden_sizes list for critic and policy")
            self._make = make_heads_config
            self._cfg_args = dict(obs_dim=q_args["obs_dim"], act_dim=q_args["act_dim"], hidden=q_args["hidden_sizes"],
                                  std_type=pi_args["std_type"], **common)
        self._max_batch = int(kwargs.get("dsact_max_batch", kwargs.get("replay_batch_size", 256)))
        self._engine = None
        self._user_seed = kwargs.get("seed", None)
        self._attachments = []
        self._register_state_dict_hook(_detach_state_dict)

    def create_action_distributions(self, logits):
        return self.policy.get_act_dist(logits)

    def _flat_groups(self):
        train = [p for n in ("q", "policy") for p in getattr(self, n).parameters()] + [self.log_alpha]
        targ = [p for n in ("q", "policy") for p in getattr(self, n + "_target").parameters()]
        return train, targ

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        if self.log_alpha.device.type == "cuda":
            self._attach(self.log_alpha.device)
        return self

    def _attach(self, device):
        eng = self._engine
        if eng is not None and eng.device != torch.device(device):
            self._engine = eng = None
        if eng is None:
            cfg = self._make(max_batch=self._max_batch, **self._cfg_args)
            eng = self._engine = CnnEngine(cfg, device, self.policy.act_high_lim, self.policy.act_low_lim)
            eng.seed(0x5DEECE66D if self._user_seed is None else int(self._user_seed))
        train, targ = self._flat_groups()
        with torch.no_grad():
            for flat, group in ((eng.params, train), (eng.targets, targ)):
                off = 0
                for p in group:
                    n = p.numel()
                    view = flat[off:off + n].view(p.shape)
                    if p.data.data_ptr() != view.data_ptr():
                        view.copy_(p.data)
                        p.data = view
                    off += n
                assert off == flat.numel(), "flat layout does not match the module"

    def engine(self, batch: int = 0) -> CnnEngine:
        if self.log_alpha.device.type != "cuda" or self._engine is None:
            raise _lib.DsactError(
                "DSAC_V1's update path runs only on the CUDA engine (libdsact.so, sm_100a); "
                "move the networks to the GPU first (`alg.networks.cuda()`). There is no CPU fallback.")
        if batch > self._max_batch:
            raise ValueError(f"batch {batch} > dsact_max_batch / replay_batch_size {self._max_batch}")
        return self._engine


def _detach_state_dict(module, state_dict, prefix, local_metadata):
    for k, v in list(state_dict.items()):
        if isinstance(v, torch.Tensor):
            state_dict[k] = v.detach().clone()
    return state_dict


class DSAC_V1:
    """DSAC (IEEE TNNLS 2021) on the B200 engine; interface of reference dsac_v1.py:56-135."""

    def __init__(self, **kwargs):
        self.networks = ApproxContainer(**kwargs)
        self.gamma = kwargs["gamma"]
        self.tau = kwargs["tau"]
        self.target_entropy = -kwargs["action_dim"]
        self.auto_alpha = kwargs["auto_alpha"]
        self.alpha = kwargs.get("alpha", 0.2)
        self.TD_bound = kwargs.get("TD_bound", 20)
        self.bound = kwargs.get("bound", True)
        self.delay_update = kwargs["delay_update"]
        self.act_dim = kwargs["action_dim"]
        self.noise_source = kwargs.get("dsact_noise", "device")
        if self.noise_source not in ("device", "reference"):
            raise ValueError("dsact_noise must be 'device' or 'reference'")

    @property
    def adjustable_parameters(self):
        return ("gamma", "tau", "auto_alpha", "alpha", "TD_bound", "bound", "delay_update")

    def _noise(self, batch: int):
        if self.noise_source == "device":
            return None
        A = self.act_dim
        eps1 = torch.empty(batch, A).normal_()   # rsample of pi(obs),         reference :147
        eps2 = torch.empty(batch, A).normal_()   # rsample of pi_target(obs2),  reference :205
        z = [torch.normal(torch.zeros(batch), torch.ones(batch)) for _ in range(3)]   # __q_evaluate x3 (:207-210, :245)
        return eps1, eps2, z[1], z[1]            # only the target critic's draw enters the update

    def local_update(self, data: Dict, iteration: int) -> dict:
        t0 = time.time()
        B = data["obs"].shape[0]
        eng = self.networks.engine(B)
        eng.step(data, iteration, self._noise(B))
        s = eng.read_stats(B)
        vals = list(s.values())
        tb = {k: vals[i] for k, i in _V1_KEYS}
        tb[tb_tags["alg_time"]] = (time.time() - t0) * 1000
        return tb

    def get_remote_update_info(self, data: Dict, iteration: int):
        raise NotImplementedError("DSAC_V1 on the B200 engine: local_update only (no gradient-message seam)")

    def remote_update(self, update_info: dict):
        raise NotImplementedError("DSAC_V1 on the B200 engine: local_update only (no gradient-message seam)")
