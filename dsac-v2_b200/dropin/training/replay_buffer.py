"""`training.replay_buffer` of the drop-in: the reference's `ReplayBuffer`
interface (reference training/replay_buffer.py:15-90) over a DEVICE-resident ring
buffer.  Transitions are staged in pinned host memory and copied to the GPU
asynchronously; `sample_batch` is a coalesced row gather on the GPU
(libdsact `dsact_replay_sample`), so no H2D copy sits on the update step.

Uniform sampling with replacement, like `np.random.randint` (reference :86).
`index_source="numpy"` draws the indices from numpy's global generator exactly as
the reference does (same seed -> same minibatch rows); the default "device" draws
them with Philox on the GPU.
"""
__all__ = ["ReplayBuffer"]

import numpy as np
import torch


class ReplayBuffer:
    _STAGE_ROWS = 4096
    _STAGES = 4

    def __init__(self, index=0, **kwargs):
        self.obsv_dim = kwargs["obsv_dim"]
        self.act_dim = kwargs["action_dim"]
        self.max_size = int(kwargs["buffer_max_size"])
        if not np.isscalar(self.act_dim):
            raise NotImplementedError("the device ring buffer stores flat action vectors")
        # image observations (CNN path, BASELINE config 5): rows hold the flattened [C*H*W] image
        self.obs_shape = None if np.isscalar(self.obsv_dim) else tuple(int(x) for x in self.obsv_dim)
        self.obs_elems = int(self.obsv_dim) if self.obs_shape is None else int(np.prod(self.obs_shape))
        if kwargs.get("additional_info"):
            raise NotImplementedError("additional_info fields are not supported by the device ring buffer")
        self.index_source = kwargs.get("dsact_index_source",
                                       "numpy" if kwargs.get("dsact_noise") == "reference" else "device")
        self.ptr, self.size = 0, 0
        self.engine = None
        self._stage = None      # pinned staging buffers
        self._events = None
        self._cur, self._fill, self._flushed = 0, 0, 0
        self._pending = []      # transitions stored before an engine was attached

    # ---- wiring -------------------------------------------------------------------
    def attach(self, engine):
        """Bind the ring storage to an engine (done by the trainer once the networks are on the GPU)."""
        eng_obs = getattr(engine, "obs_elems", None) or engine.cfg.obs_dim
        if eng_obs != self.obs_elems or engine.cfg.act_dim != self.act_dim:
            raise ValueError("replay buffer and engine disagree on obs/act dimensions")
        self.engine = engine
        engine.bind_replay(self.max_size)
        O, A = self.obs_elems, self.act_dim
        R = min(self._STAGE_ROWS if self.obs_shape is None else max(8, self._STAGE_ROWS * 400 // O), self.max_size)   # ~6 MB per staging set
        self._rows = R
        pin = lambda *s: torch.zeros(*s, dtype=torch.float32).pin_memory()
        self._stage = [dict(obs=pin(R, O), obs2=pin(R, O), act=pin(R, A), rew=pin(R), done=pin(R), logp=pin(R))
                       for _ in range(self._STAGES)]
        self._np = [{k: v.numpy() for k, v in s.items()} for s in self._stage]
        self._events = [None] * self._STAGES
        pending, self._pending = self._pending, []
        for row in pending:
            self._store_row(*row)

    def rebind(self, old, new):
        """The engine was rebuilt (larger activation arena): give the new one a ring and move the stored rows."""
        if self.engine is not old:
            return
        self.flush()
        torch.cuda.current_stream(old.device).synchronize()
        new.bind_replay(self.max_size)
        for k, v in old.replay.items():
            new.replay[k].copy_(v)
        self.engine = new

    def _require_engine(self):
        if self.engine is None:
            raise RuntimeError("ReplayBuffer is not attached to the CUDA engine: call buffer.attach(alg.networks.engine()) "
                               "(training.trainer.OffSerialTrainer does this). There is no CPU sampling path.")

    def __len__(self):
        return self.size

    def __get_RAM__(self):
        """MB of device memory holding valid transitions."""
        row_bytes = 4 * (2 * self.obs_elems + self.act_dim + 3)
        return row_bytes * self.size / 1e6

    # ---- store ----------------------------------------------------------------------
    def _store_row(self, obs, act, rew, next_obs, done, logp):
        if self._fill == self._rows:
            self.flush()
        if self._fill == 0 and self._events[self._cur] is not None:
            self._events[self._cur].synchronize()  # the async copy out of this staging buffer has finished
        s, i = self._np[self._cur], self._fill
        s["obs"][i] = obs.reshape(-1)
        s["obs2"][i] = next_obs.reshape(-1)
        s["act"][i] = act
        s["rew"][i] = rew
        s["done"][i] = done
        s["logp"][i] = logp
        self._fill += 1

    def store(self, obs, info, act, rew, next_obs, done, logp, next_info):
        row = (np.asarray(obs, dtype=np.float32), np.asarray(act, dtype=np.float32), float(rew),
               np.asarray(next_obs, dtype=np.float32), float(done), float(np.asarray(logp)))
        if self.engine is None:
            self._pending.append(row)
        else:
            self._store_row(*row)
        self.size = min(self.size + 1, self.max_size)

    def add_batch(self, samples: list):
        for sample in samples:
            self.store(*sample)

    def flush(self):
        """Enqueue the async H2D copy of the staged rows into the ring."""
        if self.engine is None or self._fill == 0:
            return
        n = self._fill
        self.engine.replay_add(self._stage[self._cur], n, self.ptr)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.engine.device))
        self._events[self._cur] = ev
        self.ptr = (self.ptr + n) % self.max_size
        self._cur = (self._cur + 1) % self._STAGES
        self._fill = 0

    # ---- full-state checkpoint ---------------------------------------
    def state_dict(self, with_data: bool = True) -> dict:
        """ptr/size (+ the valid transitions, fetched from the device ring) for an exact resume."""
        self.flush()
        out = {"ptr": self.ptr, "size": self.size, "max_size": self.max_size}
        if with_data and self.engine is not None:
            torch.cuda.current_stream(self.engine.device).synchronize()
            out["data"] = {k: v[:self.size].cpu().clone() for k, v in self.engine.replay.items()}
        return out

    def load_state_dict(self, state: dict) -> None:
        self._require_engine()
        if state["max_size"] != self.max_size:
            raise ValueError("replay capacity differs from the checkpoint")
        self.ptr, self.size, self._fill = int(state["ptr"]), int(state["size"]), 0
        if "data" in state:
            for k, v in state["data"].items():
                self.engine.replay[k][:self.size].copy_(v)

    # ---- sample -----------------------------------------------------------------------
    def sample_indices(self, batch_size: int):
        if self.index_source == "numpy":
            return torch.from_numpy(np.random.randint(0, self.size, size=batch_size))
        return None

    def sample_batch(self, batch_size: int):
        """dict of DEVICE fp32 tensors (views of the engine's minibatch arena, valid until the next sample)."""
        self._require_engine()
        if self.size == 0:
            raise ValueError("cannot sample from an empty replay buffer")
        self.flush()
        return self.engine.replay_sample(batch_size, self.size, self.sample_indices(batch_size))
