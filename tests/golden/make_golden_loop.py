#!/usr/bin/env python
"""Golden vectors for the LOOP (BASELINE.json configs[0]): the unmodified reference's OffSerialTrainer +
OffSampler + ReplayBuffer + Evaluator + DSAC_V2 on the stand-in Pendulum (tests/loop/pendulum.py), seeded, CPU.
Records the 14 deterministic tb_info scalars of every iteration and the evaluation returns.

    python tests/golden/make_golden_loop.py
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DSAC_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from dsac_v2_b200 import synth  # noqa: E402
from loop.standin import loop_kwargs  # noqa: E402

sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "gymstub"))
warnings.filterwarnings("ignore")
import torch  # noqa: E402
import utils  # noqa: E402,F401
from utils.common_utils import seed_everything  # noqa: E402
from utils.initialization import create_alg, create_buffer  # noqa: E402
from training.evaluator import create_evaluator  # noqa: E402
from training.off_sampler import create_sampler  # noqa: E402
from training.trainer import create_trainer  # noqa: E402
from make_golden import TB_KEYS  # noqa: E402

SEED, ITERS = 12345, 120

if __name__ == "__main__":
    torch.set_num_threads(4)
    save = tempfile.mkdtemp()
    os.makedirs(save + "/apprfunc", exist_ok=True)
    args = loop_kwargs(synth.reference_kwargs(synth.CONFIGS["pendulum"]), SEED, save, max_iteration=ITERS)
    seed_everything(SEED)                       # utils/init_args.py:78-79
    alg = create_alg(**args)                    # example_train/main.py:156-164
    sampler = create_sampler(**args)
    buffer = create_buffer(**args)
    evaluator = create_evaluator(**args)
    rec, rets = [], []
    inner_update, inner_eval = alg.local_update, evaluator.run_evaluation

    def local_update(data, it):
        tb = inner_update(data, it)
        rec.append([float(tb[k]) for k in TB_KEYS])
        return tb

    def run_evaluation(it):
        r = inner_eval(it)
        rets.append((it, float(r)))
        return r

    alg.local_update, evaluator.run_evaluation = local_update, run_evaluation
    trainer = create_trainer(alg, sampler, buffer, evaluator, **args)
    for _ in range(ITERS):
        trainer.step()
        trainer.iteration += 1
    out = os.path.join(HERE, "loop_pendulum.npz")
    np.savez_compressed(out, tb=np.array(rec), returns=np.array(rets), tb_keys=np.array(TB_KEYS),
                        meta=np.array([str(SEED), str(ITERS), str(buffer.size), str(sampler.get_total_sample_number())]))
    print("critic loss", rec[0][7], "->", rec[-1][7], "returns", rets, "buffer", buffer.size)
