"""Host-side pieces of the drop-in that need neither a GPU nor the CUDA library."""
import numpy as np
import pytest
import torch

from dsac_v2_b200 import _lib, dp


class _Done:
    def synchronize(self):
        pass


def test_lazy_tb_info_surfaces_data_parallel_timeouts():
    """tb_info slot 14 is the exchange status of dsact_dp_step (include/dsact.h): non-zero must raise, not log."""
    import dsac_v2
    from dsac_v2_b200.engine import STAT_KEYS
    ok = torch.arange(_lib.NUM_STATS, dtype=torch.float32)
    ok[14] = 0.0
    info = dsac_v2._LazyTbInfo(ok.clone(), _Done(), 1.5)
    assert info[STAT_KEYS[0]] == 0.0 and len(info) == len(STAT_KEYS) + 1
    bad = ok.clone()
    bad[14] = 3.0    # rank 2 never arrived
    with pytest.raises(_lib.DsactError, match="rank 2"):
        dsac_v2._LazyTbInfo(bad, _Done(), 1.5)[STAT_KEYS[0]]


@pytest.mark.parametrize("rows,world", [(37, 2), (4096, 8), (5, 8), (0, 3)])
def test_shard_rows_partitions_the_minibatch(rows, world):
    spans = [dp.shard_rows(rows, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == rows
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # contiguous, in rank order
    sizes = np.array([hi - lo for lo, hi in spans])
    assert sizes.max() - sizes.min() <= 1 and sizes.sum() == rows


def test_device_seed_mixes_run_seed_and_rank(monkeypatch):
    """every run used the same Philox key.  The engine seed is now a mix of the `seed` kwarg and the
    data-parallel rank: distinct seeds and distinct ranks give distinct 64-bit keys, the same pair reproduces."""
    import dsac_v2
    import torch.distributed as dist
    from dsac_v2_b200 import synth
    kw = synth.reference_kwargs(synth.CONFIGS["tiny"])
    seeds = {}
    for seed in (None, 0, 1, 12345):
        net = dsac_v2.ApproxContainer(**dict(kw, seed=seed))
        for rank in (0, 1, 7):
            monkeypatch.setattr(dist, "is_initialized", lambda: True)
            monkeypatch.setattr(dist, "get_rank", lambda r=rank: r)
            s = net.device_seed()
            assert 0 <= s < 2 ** 64 and s == net.device_seed()
            seeds[(seed, rank)] = s
    assert len(set(seeds.values())) == len(seeds)


def test_policy_span_of_the_trainer_covers_v2_and_v1_containers():
    """`OffSerialTrainer._find_policy_span` (the slice of the flat parameter buffer mirrored to the CPU sampler): after both
    critics of DSAC-T, after the single critic of DSAC_V1, for every policy std type."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "dsac-v2_b200", "dropin")]
    import dsac_v1
    import dsac_v2
    from dsac_v2_b200 import synth
    from training.trainer import OffSerialTrainer
    cfg = synth.CONFIGS["ragged"]
    for mod, over in ((dsac_v2, {}), (dsac_v2, {"policy_std_type": "mlp_separated"}), (dsac_v2, {"policy_std_type": "parameter"}),
                      (dsac_v1, {"algorithm": "DSAC_V1"})):
        net = mod.ApproxContainer(**synth.reference_kwargs(cfg, replay_batch_size=8, **over))

        class Probe:
            networks = net
        lo, hi = OffSerialTrainer._find_policy_span(Probe)
        train, _ = net._flat_groups()
        flat_names = []
        for name in (("q1", "q2") if hasattr(net, "q1") else ("q",)) + ("policy",):
            flat_names += [name] * sum(p.numel() for p in getattr(net, name).parameters())
        assert flat_names[lo:hi] == ["policy"] * (hi - lo) and "policy" not in flat_names[:lo]
        assert hi == sum(p.numel() for p in train) - 1     # log_alpha is the last element
