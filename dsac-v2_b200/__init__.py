"""dsac-v2_b200 — B200-native engine for ONE path of Jingliang-Duan/DSAC-v2:
the per-step DSAC-T update over a replay minibatch (`DSAC_V2.local_update`,
reference dsac_v2.py:102-105) and the replay gather that feeds it.

Layout
  csrc/        CUDA kernels + C ABI (include/dsact.h) -> libdsact.so
  _lib.py      ctypes binding          engine.py   torch-owned buffers around one handle
  synth.py     deterministic synthetic inputs (tests / bench / golden generator)
  dropin/      host-side mirror of the reference interface for this path; put this
               directory on sys.path ahead of a DSAC-v2 checkout and `dsac_v2`,
               `networks.mlp`, `training.replay_buffer`, `training.trainer` resolve here.
"""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
DROPIN_DIR = os.path.join(PACKAGE_DIR, "dropin")


def enable_dropin() -> str:
    """Put the drop-in modules first on sys.path (idempotent); returns the directory."""
    import sys

    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR
