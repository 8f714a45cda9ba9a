"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and the size/validation entry points (which need no GPU) behave."""
import ctypes as C
import os
import re

import pytest

from dsac_v2_b200 import _lib, synth
from dsac_v2_b200.engine import make_config, query_layout

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    header = open(os.path.join(REPO, "include", "dsact.h")).read()
    declared = set(re.findall(r"\b(dsact_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dsact.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.dsact_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header_layout():
    # 5 + 12 + 8 int32 = 25 int32 (100 B, padded to 104 for the doubles) + 12 doubles
    assert C.sizeof(_lib.Config) == 104 + 12 * 8
    assert C.sizeof(_lib.Batch) == 5 * 8 + 8 + 8
    assert C.sizeof(_lib.Buffers) == 9 * 8


@pytest.mark.parametrize("name", list(synth.CONFIGS))
def test_layout_matches_network_shapes(name):
    cfg = synth.CONFIGS[name]
    q, pi = synth.net_shapes(cfg["obs_dim"], cfg["act_dim"], cfg["hidden"])
    count = lambda s: sum(s[j] * s[j + 1] + s[j + 1] for j in range(len(s) - 1))
    lay = query_layout(make_config(cfg["obs_dim"], cfg["act_dim"], cfg["hidden"], cfg["hidden"], max_batch=256))
    assert lay.n_q == count(q) and lay.n_pi == count(pi)
    assert lay.n_params == 2 * count(q) + count(pi) + 1
    assert lay.n_targets == lay.n_params - 1
    assert lay.workspace_bytes > 0 and lay.workspace_bytes % 256 == 0


def test_humanoid_parameter_count_is_the_surveyed_one():
    cfg = synth.CONFIGS["humanoid"]
    lay = query_layout(make_config(376, 17, cfg["hidden"], cfg["hidden"], max_batch=4096))
    assert (lay.n_q, lay.n_pi, lay.n_params) == (232962, 236834, 702759)


def test_invalid_configs_are_rejected_with_a_message():
    lib = _lib.load()
    out = _lib.Layout()
    for mutate in (lambda c: setattr(c, "obs_dim", 0), lambda c: setattr(c, "n_hidden_q", 0),
                   lambda c: setattr(c, "delay_update", 0), lambda c: setattr(c, "abi_version", 99),
                   lambda c: setattr(c, "act_q", 42), lambda c: setattr(c, "max_batch", 0)):
        c = make_config(5, 2, [8], [8], max_batch=4)
        mutate(c)
        assert lib.dsact_query_layout(C.byref(c), C.byref(out)) == -1
        assert lib.dsact_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a CUDA-less host")
    from dsac_v2_b200.engine import Engine
    with pytest.raises(_lib.DsactError):
        Engine(make_config(5, 2, [8], [8], max_batch=4), torch.device("cuda", 0), torch.ones(2), -torch.ones(2))
    import dsac_v2
    alg = dsac_v2.DSAC_V2(**synth.reference_kwargs(synth.CONFIGS["tiny"]))
    with pytest.raises(_lib.DsactError):
        alg.local_update({"obs": torch.zeros(4, 5)}, 0)


@pytest.mark.parametrize("variant", ["cnn_type2", "cnn_type1", "mlp_separated", "parameter", "v1_mlp"])
def test_head_wise_layouts_match_the_dropin_modules(variant):
    """The flat layout the head-wise engine reports (`dsact_cnn_query_layout`, no GPU needed) and the state_dict schema
    `CnnEngine._schema` walks are those of the drop-in modules' own parameter order, for every variant it serves."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "dsac-v2_b200", "dropin")]
    import dsac_v1
    import dsac_v2
    from dsac_v2_b200.engine_cnn import CnnEngine
    from dsac_v2_b200._lib import Layout
    if variant.startswith("cnn"):
        cfg = synth.CNN_CONFIGS["small_t1" if variant != "cnn_type2" else "carracing"]
        kw = synth.cnn_reference_kwargs(cfg, replay_batch_size=4)
    else:
        cfg = synth.CONFIGS["ragged"]
        over = {"algorithm": "DSAC_V1"} if variant == "v1_mlp" else {"policy_std_type": variant}
        kw = synth.reference_kwargs(cfg, replay_batch_size=4, **over)
    net = (dsac_v1 if variant.startswith("v1") else dsac_v2).ApproxContainer(**kw)
    make = net._make if variant.startswith("v1") else None
    if make is None:
        from dsac_v2_b200.engine_cnn import make_cnn_config, make_heads_config
        make = make_heads_config if net._heads_std else make_cnn_config
    c = make(max_batch=4, **net._cfg_args)
    lay = Layout()
    assert _lib.load().dsact_cnn_query_layout(C.byref(c), C.byref(lay)) == 0
    train, targ = net._flat_groups()
    assert lay.n_params == sum(p.numel() for p in train) and lay.n_targets == sum(p.numel() for p in targ)

    class Probe:   # _schema only reads the config
        cfg = c
    schema, n = CnnEngine._schema(Probe)
    assert n == lay.n_targets
    names = [k for k, p in net.named_parameters() if p.requires_grad and k != "log_alpha"]
    assert [e[0] for e in schema] == names
    sizes = dict(net.named_parameters())
    assert all(tuple(sizes[e[0]].shape) == tuple(e[4]) and sizes[e[0]].numel() == e[3] for e in schema)
    assert [e[1] for e in schema] == [k for k, p in net.named_parameters() if not p.requires_grad]
