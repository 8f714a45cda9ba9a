"""World-size-2 CPU test (gloo) of the data-parallel seam: the collective sequence of
`dsac_v2_b200.dp.data_parallel_gradients` driven by a CPU stand-in engine (the oracle's shard
arithmetic) must reproduce the single-process full-batch update."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    """grad_phase1 / grad_phase2 / state / grads with the layout of include/dsact.h, on the CPU oracle."""

    def __init__(self, orc):
        from dsac_v2_b200 import _lib
        self.orc, self._lib = orc, _lib
        self.state = torch.zeros(64)
        self.grads = None
        self._pending = None

    def grad_phase1(self, data, noise):
        # forward up to the critic-std sums: run the oracle with a hook that records the local sums
        self._pending = (data, noise)
        rec = {}

        def hook(sums):
            rec["s"] = [float(x) for x in sums]
            raise StopIteration

        saved = list(self.orc.mean_std)
        try:
            self.orc.compute_gradients(data, noise, global_batch=1, std_sum_hook=hook)
        except StopIteration:
            pass
        self.orc.mean_std = saved
        self.state[self._lib.STATE_STDSUM:self._lib.STATE_STDSUM + 2] = torch.tensor(rec["s"])

    def grad_phase2(self, global_rows):
        data, noise = self._pending
        reduced = self.state[self._lib.STATE_STDSUM:self._lib.STATE_STDSUM + 2].clone()
        self.tb = self.orc.compute_gradients(data, noise, global_batch=global_rows,
                                             std_sum_hook=lambda sums: [reduced[0], reduced[1]])
        g = self.orc.grads
        flat = [t.reshape(-1) for n in ("q1", "q2", "policy") for t in g[n]] + [g["log_alpha"][0].reshape(1)]
        self.grads = torch.cat(flat).clone()

    def scatter_grads(self):
        off = 0
        for n in ("q1", "q2", "policy"):
            for i, t in enumerate(self.orc.grads[n]):
                self.orc.grads[n][i] = self.grads[off:off + t.numel()].view_as(t).clone()
                off += t.numel()
        self.orc.grads["log_alpha"] = [self.grads[off].clone()]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsac_v2_b200 import dp, synth
    from oracle.dsact_oracle import from_config
    cfg, B = synth.CONFIGS["ragged"], 37
    eng = OracleShardEngine(from_config(cfg, synth.make_weights(cfg), **synth.HYPER))
    d, w = dp.world()
    assert w == world
    for it in range(4):
        full, noise = synth.make_batch(cfg, B, it), synth.make_noise(cfg, B, it)
        lo, hi = dp.shard_rows(B, rank, world)  # ragged: 19 + 18 rows
        shard = {k: torch.from_numpy(v[lo:hi]) for k, v in full.items()}
        nshard = [torch.from_numpy(n[lo:hi]) for n in noise]
        rows = dp.global_rows(d, hi - lo, torch.device("cpu"))
        assert rows == B
        dp.data_parallel_gradients(eng, shard, nshard, d, hi - lo, rows)
        eng.scatter_grads()
        eng.orc.apply(it)
    sd = {k: v.numpy() for k, v in eng.orc.state_dict().items()}
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **sd)
    dist.destroy_process_group()


def test_two_rank_shards_equal_full_batch(tmp_path):
    from dsac_v2_b200 import synth
    from oracle.dsact_oracle import from_config
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cfg, B = synth.CONFIGS["ragged"], 37
    ref = from_config(cfg, synth.make_weights(cfg), **synth.HYPER)
    for it in range(4):
        ref.update(synth.make_batch(cfg, B, it), synth.make_noise(cfg, B, it), it)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k, v in ref.state_dict().items():
        np.testing.assert_array_equal(r0[k], r1[k], err_msg=f"replicas diverged: {k}")
        np.testing.assert_allclose(r0[k], v.numpy(), rtol=2e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("n,world", [(37, 2), (4096, 8), (5, 8), (0, 3)])
def test_shard_rows_partition(n, world):
    from dsac_v2_b200 import dp
    spans = [dp.shard_rows(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


class _PeerStub:
    """dp_export / dp_connect of the CUDA engine, scripted: which rank fails where."""

    def __init__(self, rank, fail_export_on=None, fail_connect_on=None):
        self.rank, self.fail_export_on, self.fail_connect_on = rank, fail_export_on, fail_connect_on
        self.dp_world, self.connected = 0, None

    def dp_export(self):
        from dsac_v2_b200 import _lib
        if self.rank == self.fail_export_on:
            raise _lib.DsactError("cudaIpcGetMemHandle: not supported")
        return bytes([self.rank]) * _lib.IPC_HANDLE_BYTES

    def dp_connect(self, rank, handles):
        from dsac_v2_b200 import _lib
        if rank == self.fail_connect_on:
            raise _lib.DsactError("cudaIpcOpenMemHandle: peer access denied")
        self.connected, self.dp_world = list(handles), len(handles)


def _peer_worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsac_v2_b200 import _lib, dp
    results = []
    ok = _PeerStub(rank)
    results.append(dp.connect_peers(ok, dist))                       # every rank fine: handles in rank order everywhere
    assert ok.connected == [bytes([r]) * _lib.IPC_HANDLE_BYTES for r in range(world)] and ok.dp_world == world
    bad_export = _PeerStub(rank, fail_export_on=1)
    results.append(dp.connect_peers(bad_export, dist))               # one rank cannot export: nobody connects
    assert bad_export.connected is None and bad_export.dp_world == 0
    bad_connect = _PeerStub(rank, fail_connect_on=0)
    results.append(dp.connect_peers(bad_connect, dist))              # one rank cannot map a peer: all fall back alike
    assert bad_connect.dp_world == 0
    np.save(os.path.join(out_dir, f"peer{rank}.npy"), np.array(results))
    dist.destroy_process_group()


def test_peer_setup_is_all_or_nothing(tmp_path):
    """`dp.connect_peers` must return the same answer on every rank, whichever rank fails to export or map a buffer
    (a split decision would leave some ranks spinning on peers that took the NCCL path)."""
    port = 29800 + os.getpid() % 1000
    mp.spawn(_peer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "peer0.npy"), np.load(tmp_path / "peer1.npy")
    assert r0.tolist() == r1.tolist() == [True, False, False]
