"""Python owner of one libdsact handle: allocates the flat device buffers with
torch (the library only borrows pointers), and exposes the update path.

This is plumbing around the C ABI (include/dsact.h); all arithmetic of the
path runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import Batch, Buffers, Config, Layout, Noise, Replay, check

STAT_KEYS = [
    "DSAC2/critic_avg_q1-RL iter",
    "DSAC2/critic_avg_q2-RL iter",
    "DSAC2/critic_avg_std1-RL iter",
    "DSAC2/critic_avg_std2-RL iter",
    "DSAC2/critic_avg_min_std1-RL iter",
    "DSAC2/critic_avg_min_std2-RL iter",
    "Loss/Actor loss-RL iter",
    "Loss/Critic loss-RL iter",
    "DSAC2/policy_mean-RL iter",
    "DSAC2/policy_std-RL iter",
    "DSAC2/entropy-RL iter",
    "DSAC2/alpha-RL iter",
    "DSAC2/mean_std1",
    "DSAC2/mean_std2",
]


def make_config(obs_dim: int, act_dim: int, hidden_q: Sequence[int], hidden_pi: Sequence[int], *, max_batch: int,
                act_q: str = "gelu", act_pi: str = "gelu", gamma=0.99, tau=0.005, tau_b=None, delay_update=2,
                auto_alpha=True, alpha=0.2, lr_q=1e-4, lr_pi=1e-4, lr_alpha=3e-4, min_log_std=-20.0,
                max_log_std=0.5, gemm_mode="fp32", use_graph=True, act_dist="TanhGaussDistribution") -> Config:
    if len(hidden_q) > _lib.MAX_HIDDEN or len(hidden_pi) > _lib.MAX_HIDDEN:
        raise ValueError(f"at most {_lib.MAX_HIDDEN} hidden layers")
    for name in (act_q, act_pi):
        if name not in _lib.ACTIVATIONS:
            raise ValueError(f"unsupported activation {name!r}")
    c = Config()
    c.abi_version = _lib.ABI_VERSION
    c.obs_dim, c.act_dim = int(obs_dim), int(act_dim)
    c.n_hidden_q, c.n_hidden_pi = len(hidden_q), len(hidden_pi)
    for j, v in enumerate(hidden_q):
        c.hidden_q[j] = int(v)
    for j, v in enumerate(hidden_pi):
        c.hidden_pi[j] = int(v)
    c.act_q, c.act_pi = _lib.ACTIVATIONS[act_q], _lib.ACTIVATIONS[act_pi]
    c.max_batch = int(max_batch)
    c.auto_alpha, c.delay_update = int(bool(auto_alpha)), int(delay_update)
    c.gemm_mode, c.use_graph = _lib.GEMM_MODES[gemm_mode], int(bool(use_graph))
    c.act_dist = _lib.ACT_DISTS[act_dist]
    c.gamma, c.tau = float(gamma), float(tau)
    c.tau_b = float(tau if tau_b is None else tau_b)
    c.alpha_fixed = float(alpha)
    c.lr_q, c.lr_pi, c.lr_alpha = float(lr_q), float(lr_pi), float(lr_alpha)
    c.min_log_std, c.max_log_std = float(min_log_std), float(max_log_std)
    c.adam_beta1, c.adam_beta2, c.adam_eps = 0.9, 0.999, 1e-8
    return c


def query_layout(cfg: Config) -> Layout:
    out = Layout()
    check(_lib.load().dsact_query_layout(C.byref(cfg), C.byref(out)))
    return out


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


class Engine:
    """One handle bound to flat torch-owned buffers on one CUDA device."""

    def __init__(self, cfg: Config, device: torch.device, act_high: torch.Tensor, act_low: torch.Tensor):
        if not torch.cuda.is_available():
            raise _lib.DsactError("the DSAC-T update engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DsactError(f"engine device must be CUDA, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.layout = query_layout(cfg)
        L = self.layout
        with torch.cuda.device(self.device):
            z = lambda n: torch.zeros(int(n), dtype=torch.float32, device=self.device)
            self.params, self.targets = z(L.n_params), z(L.n_targets)
            self.grads, self.adam_m, self.adam_v = z(L.n_params), z(L.n_params), z(L.n_params)
            self.state = z(L.state_floats)
            self.workspace = torch.zeros(int(L.workspace_bytes) // 4 + 64, dtype=torch.float32, device=self.device)
            self.act_high = _f32c(torch.as_tensor(act_high).reshape(-1), self.device).clone()
            self.act_low = _f32c(torch.as_tensor(act_low).reshape(-1), self.device).clone()
            h = C.c_void_p()
            check(self.lib.dsact_create(C.byref(cfg), self.device.index, C.byref(h)))
            self.h = h
            self._stats_host = torch.zeros(_lib.NUM_STATS, dtype=torch.float32).pin_memory()
            self._bind()
            check(self.lib.dsact_set_carry(self.h, -1.0, -1.0, 0, 0, self._stream()))
        self.replay = None
        self.dp_world = 0          # > 1 once dp_connect has mapped the peers
        self._seed = 0x5DEECE66D   # the library's default (csrc/engine.cu)
        self._staged = False
        self._keep_host = [None, None]
        self._keep = None  # tensors referenced by the last enqueued call

    def _bind(self):
        ws = self.workspace
        off = (-ws.data_ptr() % 256) // 4  # 256-byte aligned view
        self._ws_view = ws[off:]
        b = Buffers(self.params.data_ptr(), self.targets.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                    self.adam_v.data_ptr(), self.act_high.data_ptr(), self.act_low.data_ptr(), self.state.data_ptr(),
                    self._ws_view.data_ptr())
        check(self.lib.dsact_bind(self.h, C.byref(b)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsact_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- argument marshalling ------------------------------------------------
    def _host_batch(self, data) -> Batch:
        """ctypes view of a HOST minibatch (contiguous fp32 CPU tensors; pinned ones copy at full PCIe rate)."""
        t = {}
        for k in ("obs", "act", "rew", "obs2", "done"):
            v = data[k]
            if v.dtype != torch.float32 or not v.is_contiguous():
                v = v.to(torch.float32).contiguous()
            t[k] = v
        B = t["obs"].shape[0]
        O, A = self.cfg.obs_dim, self.cfg.act_dim
        if B > self.cfg.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.cfg.max_batch}")
        if t["obs"].shape != (B, O) or t["obs2"].shape != (B, O) or t["act"].shape != (B, A) \
                or t["rew"].numel() != B or t["done"].numel() != B:
            raise ValueError("minibatch shapes do not match the configured obs_dim/act_dim")
        # the async copies read these tensors after the call returns: keep the last two sets alive
        self._keep_host = (getattr(self, "_keep_host", None) or [None, None])[1:] + [t]
        return Batch(t["obs"].data_ptr(), t["act"].data_ptr(), t["rew"].data_ptr(), t["obs2"].data_ptr(),
                     t["done"].data_ptr(), B, None)

    def _stage_in(self, data) -> Batch:
        """Host minibatch -> one of the library's two device staging sets, copied on its private copy stream so that the
        H2D transfer of call k+1 overlaps the kernels of call k (`dsact_stage_host`, include/dsact.h)."""
        hb = self._host_batch(data)
        dev = Batch()
        check(self.lib.dsact_stage_host(self.h, C.byref(hb), C.byref(dev), self._stream()))
        self._staged = True
        return dev

    def _mark_staged_done(self):
        """The kernels reading the current staging set have been enqueued: the set may be refilled after them."""
        if getattr(self, "_staged", False):
            check(self.lib.dsact_stage_release(self.h, self._stream()))
            self._staged = False

    def _batch(self, data: Dict[str, torch.Tensor]) -> Batch:
        if data["obs"].device.type == "cpu":
            b = self._stage_in(data)
            self._keep = None
            return b
        t = {k: _f32c(data[k], self.device) for k in ("obs", "act", "rew", "obs2", "done")}
        B = t["obs"].shape[0]
        O, A = self.cfg.obs_dim, self.cfg.act_dim
        if t["obs"].shape != (B, O) or t["obs2"].shape != (B, O) or t["act"].shape != (B, A) \
                or t["rew"].numel() != B or t["done"].numel() != B:
            raise ValueError("minibatch shapes do not match the configured obs_dim/act_dim")
        self._keep = t
        return Batch(t["obs"].data_ptr(), t["act"].data_ptr(), t["rew"].data_ptr(), t["obs2"].data_ptr(),
                     t["done"].data_ptr(), B, None)

    def _noise_slots(self, B):
        O, A, r64 = self.cfg.obs_dim, self.cfg.act_dim, lambda n: (n + 63) // 64 * 64
        mb = self.cfg.max_batch
        off = 2 * r64(mb * O) + r64(mb * A) + 3 * r64(mb) + r64(2 * mb)  # obs obs2 act rew done logp idx
        out = []
        for n, shape in ((A, (B, A)), (A, (B, A)), (1, (B,)), (1, (B,))):
            out.append(self._ws_view[off:off + B * n].view(shape))
            off += r64(mb * n)
        return out

    def _noise(self, noise, B):
        if noise is None:
            return None, None
        noise = [torch.as_tensor(x) for x in noise]
        if all(x.device.type == "cpu" for x in noise):  # stage into the arena's noise slots: stable pointers
            slots = self._noise_slots(B)
            for dst, src in zip(slots, noise):
                dst.copy_(src.reshape(dst.shape), non_blocking=True)
            noise = slots
        eps1, eps2, z3, z4 = (_f32c(x, self.device) for x in noise)
        A = self.cfg.act_dim
        if eps1.shape != (B, A) or eps2.shape != (B, A) or z3.numel() != B or z4.numel() != B:
            raise ValueError("noise shapes must be eps1/eps2 [B,A], z3/z4 [B]")
        keep = (eps1, eps2, z3, z4)
        return C.byref(Noise(eps1.data_ptr(), eps2.data_ptr(), z3.data_ptr(), z4.data_ptr())), keep

    # ---- the path ---------------------------------------------------------------
    def step(self, data, iteration: int, noise=None):
        """DSAC_V2.local_update (reference dsac_v2.py:102-105) on device tensors."""
        with torch.cuda.device(self.device):
            if data["obs"].device.type == "cpu":   # the reference-facing call with a host minibatch: one C call
                b = self._host_batch(data)
                n, keep = self._noise(noise, b.batch)
                self._keep_noise = keep
                check(self.lib.dsact_step_host(self.h, C.byref(b), n, int(iteration), self._stream()))
            else:
                b = self._batch(data)
                n, keep = self._noise(noise, b.batch)
                self._keep_noise = keep
                check(self.lib.dsact_step(self.h, C.byref(b), n, int(iteration), self._stream()))
        self.last_batch = b.batch

    def profile_step(self, data, iteration: int, noise=None) -> dict:
        """One eager step with per-launch CUDA events (bench.py's roofline leg)."""
        out = _lib.Profile()
        with torch.cuda.device(self.device):
            b = self._batch(data)
            n, keep = self._noise(noise, b.batch)
            check(self.lib.dsact_profile_step(self.h, C.byref(b), n, int(iteration), self._stream(), C.byref(out)))
            self._mark_staged_done()
        self.last_batch = b.batch
        names = ("other", "gemm_fwd", "gemm_dgrad", "gemm_wgrad")
        return {"total_ms": out.total_ms,
                **{k: {"ms": out.ms[i], "flops": out.flops[i], "launches": out.launches[i]} for i, k in enumerate(names)}}

    def compute_grads(self, data, noise=None):
        with torch.cuda.device(self.device):
            b = self._batch(data)
            n, keep = self._noise(noise, b.batch)
            self._keep_noise = keep
            check(self.lib.dsact_compute_grads(self.h, C.byref(b), n, self._stream()))
            self._mark_staged_done()
        self.last_batch = b.batch

    def grad_phase1(self, data, noise=None):
        with torch.cuda.device(self.device):
            b = self._batch(data)
            n, keep = self._noise(noise, b.batch)
            self._keep_noise = keep
            check(self.lib.dsact_grad_phase1(self.h, C.byref(b), n, self._stream()))
        self.last_batch = b.batch

    def grad_phase2(self, global_batch: int):
        with torch.cuda.device(self.device):
            check(self.lib.dsact_grad_phase2(self.h, int(global_batch), self._stream()))
            self._mark_staged_done()

    def apply(self, iteration: int):
        with torch.cuda.device(self.device):
            check(self.lib.dsact_apply(self.h, int(iteration), self._stream()))

    def read_stats_async(self, global_batch: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Enqueue finalisation + D2H of the 16 tb_info floats; caller synchronises before reading."""
        out = self._stats_host if out is None else out
        with torch.cuda.device(self.device):
            check(self.lib.dsact_read_stats(self.h, int(global_batch or self.last_batch), out.data_ptr(), self._stream()))
        return out

    def read_stats(self, global_batch: Optional[int] = None) -> Dict[str, float]:
        out = self.read_stats_async(global_batch)
        torch.cuda.current_stream(self.device).synchronize()
        if float(out[14]) != 0.0:   # include/dsact.h: slot 14 = 1 + rank of a peer that never arrived (dsact_dp_step)
            raise _lib.DsactError(f"data-parallel exchange timed out waiting for rank {int(out[14]) - 1}")
        return {k: float(out[i]) for i, k in enumerate(STAT_KEYS)}

    def set_carry(self, mean_std1=-1.0, mean_std2=-1.0, adam_steps_q=0, adam_steps_pi=0):
        with torch.cuda.device(self.device):
            check(self.lib.dsact_set_carry(self.h, float(mean_std1), float(mean_std2), int(adam_steps_q),
                                           int(adam_steps_pi), self._stream()))

    def seed(self, seed: int):
        self._seed = int(seed) & (2 ** 64 - 1)
        check(self.lib.dsact_seed(self.h, self._seed))

    # ---- replay ring buffer -------------------------------------------------------
    def bind_replay(self, capacity: int):
        O, A = self.cfg.obs_dim, self.cfg.act_dim
        with torch.cuda.device(self.device):
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
            self.replay = dict(obs=z(capacity, O), obs2=z(capacity, O), act=z(capacity, A), rew=z(capacity),
                               done=z(capacity), logp=z(capacity))
            r = self.replay
            rb = Replay(r["obs"].data_ptr(), r["obs2"].data_ptr(), r["act"].data_ptr(), r["rew"].data_ptr(),
                        r["done"].data_ptr(), r["logp"].data_ptr(), int(capacity))
            check(self.lib.dsact_replay_bind(self.h, C.byref(rb)))
        self.capacity = int(capacity)

    def replay_add(self, staging: Dict[str, torch.Tensor], n: int, ptr: int):
        """Rows of contiguous fp32 staging tensors (pinned host or device) -> ring rows (ptr+i) % capacity."""
        s = staging
        with torch.cuda.device(self.device):
            check(self.lib.dsact_replay_add(self.h, s["obs"].data_ptr(), s["obs2"].data_ptr(), s["act"].data_ptr(),
                                            s["rew"].data_ptr(), s["done"].data_ptr(), s["logp"].data_ptr(),
                                            int(n), int(ptr), self._stream()))

    def arena_batch(self, b: Batch) -> Dict[str, torch.Tensor]:
        """Torch views of the engine's gathered-minibatch arena (no copies)."""
        O, A, B = self.cfg.obs_dim, self.cfg.act_dim, b.batch
        base = self._ws_view.data_ptr()

        def view(ptr, n, shape):
            off = (ptr - base) // 4
            return self._ws_view[off:off + n].view(shape)

        return {"obs": view(b.obs, B * O, (B, O)), "act": view(b.act, B * A, (B, A)), "rew": view(b.rew, B, (B,)),
                "obs2": view(b.obs2, B * O, (B, O)), "done": view(b.done, B, (B,)), "logp": view(b.logp, B, (B,))}

    def replay_sample(self, batch: int, size: int, idx: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        out = Batch()
        with torch.cuda.device(self.device):
            if idx is not None:
                idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
                self._keep_idx = idx
            check(self.lib.dsact_replay_sample(self.h, int(batch), int(size), _ptr(idx), C.byref(out), self._stream()))
        return self.arena_batch(out)

    def replay_step(self, batch: int, size: int, iteration: int, idx: Optional[torch.Tensor] = None, noise=None):
        with torch.cuda.device(self.device):
            if idx is not None:
                idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
                self._keep_idx = idx
            n, keep = self._noise(noise, batch)
            self._keep_noise = keep
            check(self.lib.dsact_replay_step(self.h, int(batch), int(size), _ptr(idx), n, int(iteration), self._stream()))
        self.last_batch = int(batch)

    # ---- data-parallel replicas over NVLink peer memory (include/dsact.h) ---------------------
    def dp_export(self) -> bytes:
        """Allocate this rank's exchange buffer; its CUDA IPC handle (to be handed to every other rank)."""
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        n = C.c_int64(0)
        with torch.cuda.device(self.device):
            check(self.lib.dsact_dp_export(self.h, buf, C.byref(n)))
        return buf.raw

    def dp_connect(self, rank: int, handles: Sequence[bytes]):
        """Map every rank's exchange buffer (`handles` in rank order, one per rank including this one)."""
        blob = b"".join(handles)
        if len(blob) != _lib.IPC_HANDLE_BYTES * len(handles):
            raise ValueError("malformed IPC handle list")
        with torch.cuda.device(self.device):
            check(self.lib.dsact_dp_connect(self.h, int(rank), len(handles), blob))
        self.dp_world = len(handles)

    def dp_step(self, data, iteration: int, global_batch: int, noise=None):
        """dsact_step on this rank's shard with the exchanges done in-kernel over peer memory."""
        with torch.cuda.device(self.device):
            b = self._batch(data)
            n, keep = self._noise(noise, b.batch)
            self._keep_noise = keep
            check(self.lib.dsact_dp_step(self.h, C.byref(b), n, int(global_batch), int(iteration), self._stream()))
            self._mark_staged_done()
        self.last_batch = b.batch

    def dp_replay_step(self, batch: int, size: int, iteration: int, global_batch: int, idx: Optional[torch.Tensor] = None,
                       noise=None):
        with torch.cuda.device(self.device):
            if idx is not None:
                idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
                self._keep_idx = idx
            n, keep = self._noise(noise, batch)
            self._keep_noise = keep
            check(self.lib.dsact_dp_replay_step(self.h, int(batch), int(size), _ptr(idx), n, int(global_batch), int(iteration),
                                                self._stream()))
        self.last_batch = int(batch)

    # ---- weights in the reference's state_dict schema -----------------------------------
    def _schema(self):
        """[(key, flat name, offset, shape)] for every tensor of the flat layout (include/dsact.h)."""
        c = self.cfg
        q_sizes = [c.obs_dim + c.act_dim] + [c.hidden_q[j] for j in range(c.n_hidden_q)] + [2]
        pi_sizes = [c.obs_dim] + [c.hidden_pi[j] for j in range(c.n_hidden_pi)] + [2 * c.act_dim]
        out, off = [], 0
        for net, inner, sizes in (("q1", "q", q_sizes), ("q2", "q", q_sizes), ("policy", "policy", pi_sizes)):
            for j in range(len(sizes) - 1):
                for leaf, shape in (("weight", (sizes[j + 1], sizes[j])), ("bias", (sizes[j + 1],))):
                    n = 1
                    for d in shape:
                        n *= d
                    out.append((f"{net}.{inner}.{2 * j}.{leaf}", f"{net}_target.{inner}.{2 * j}.{leaf}", off, n, shape))
                    off += n
        return out, off

    def load_weights(self, weights: dict):
        """Fill params/targets from a dict keyed like the reference's state_dict."""
        schema, n = self._schema()
        with torch.no_grad():
            for key, tkey, off, cnt, shape in schema:
                self.params[off:off + cnt].copy_(torch.as_tensor(weights[key]).reshape(-1))
                self.targets[off:off + cnt].copy_(torch.as_tensor(weights.get(tkey, weights[key])).reshape(-1))
            self.params[n] = float(weights.get("log_alpha", 1.0))

    def export_weights(self, grads: bool = False) -> dict:
        schema, n = self._schema()
        src = self.grads if grads else self.params
        host, thost = src.detach().cpu(), self.targets.detach().cpu()
        out = {"log_alpha": host[n].clone()}
        for key, tkey, off, cnt, shape in schema:
            out[key] = host[off:off + cnt].view(shape).clone()
            if not grads:
                out[tkey] = thost[off:off + cnt].view(shape).clone()
        return out

    # ---- test hooks ----------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.dsact_launch_count(self.h))

    def last_call_launches(self) -> int:
        return int(self.lib.dsact_last_call_launches(self.h))

    def test_gemm(self, variant: int, A: torch.Tensor, B: torch.Tensor, bias: Optional[torch.Tensor], C_out: torch.Tensor,
                  M: int, N: int, K: int):
        with torch.cuda.device(self.device):
            check(self.lib.dsact_test_gemm(self.h, variant, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                           _ptr(bias), C_out.data_ptr(), C_out.stride(0), M, N, K, self._stream()))
