"""Per-CTA, per-layer timeline of the fused layer-chain kernels (DSACT_TC_DEBUG=1): eager steps at the bench shape.
usage: python tools/chain_timeline.py [config] [batch] [gemm_mode] [activation]"""
import os
import sys

os.environ["DSACT_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dsac_v2_b200 import synth  # noqa: E402
from dsac_v2_b200.engine import Engine, make_config  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "humanoid"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
act = sys.argv[4] if len(sys.argv) > 4 else "gelu"
c = make_config(cfg["obs_dim"], cfg["act_dim"], cfg["hidden"], cfg["hidden"], max_batch=B, gemm_mode=mode, use_graph=False, act_q=act, act_pi=act)
lim = torch.full((cfg["act_dim"],), cfg["act_lim"])
eng = Engine(c, torch.device("cuda", 0), lim, -lim)
eng.load_weights(synth.make_weights(cfg))
batch = {k: torch.from_numpy(v).cuda() for k, v in synth.make_batch(cfg, B, 0).items()}
for it in range(3):
    print(f"==== step {it} ====", file=sys.stderr, flush=True)
    eng.step(batch, it)
torch.cuda.synchronize()
