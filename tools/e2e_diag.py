#!/usr/bin/env python
"""Where does the end-to-end (host minibatch -> local_update) time go?  Prints the box's NUMA layout, the pinned
H2D bandwidth of a 12.6 MB transfer from every NUMA node, and the local_update loop rate under a few host settings.

    python tools/e2e_diag.py [--steps 300]
"""
import argparse
import glob
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "dsac-v2_b200", "dropin"))
import torch  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def node_cpus():
    out = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(p.rsplit("node", 1)[1])
        cpus = set()
        for part in open(os.path.join(p, "cpulist")).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        out[n] = cpus
    return out


def h2d_gbs(nbytes=12_648_448, reps=40):
    src = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    src.fill_(1.0)
    dst = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    print("== topology")
    print(sh("nvidia-smi topo -m | head -20"))
    print(sh("lscpu | grep -i -E 'numa|model name|socket|^CPU\\(s\\)'"))
    bdf = sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0").lower()
    bdf = bdf[4:] if len(bdf) > 12 else bdf
    print("gpu0 bdf", bdf, "numa_node", sh(f"cat /sys/bus/pci/devices/{bdf}/numa_node"))
    aff0 = os.sched_getaffinity(0)
    print("affinity:", len(aff0), "cpus", sorted(aff0)[:4], "...", "torch threads", torch.get_num_threads())
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    print("== pinned H2D bandwidth of 12.6 MB copies")
    print(f"default affinity: {h2d_gbs():.1f} GB/s")
    nodes = node_cpus()
    for n, cpus in nodes.items():
        use = cpus & aff0
        if not use:
            continue
        os.sched_setaffinity(0, use)
        print(f"node {n} ({len(use)} cpus): {h2d_gbs():.1f} GB/s")
    os.sched_setaffinity(0, aff0)

    print("== local_update loop (host pinned minibatches)")
    import dsac_v2
    from dsac_v2_b200 import synth
    cfg = synth.CONFIGS["humanoid"]
    B, O, A = 4096, cfg["obs_dim"], cfg["act_dim"]
    kw = synth.reference_kwargs(cfg, replay_batch_size=B, dsact_gemm="bf16x3", buffer_max_size=1000, additional_info={})
    alg = dsac_v2.DSAC_V2(**kw)
    alg.networks.cuda()

    def loop(tag, read=True):
        hg = torch.Generator().manual_seed(7)
        ring = [{"obs": torch.randn(B, O, generator=hg).pin_memory(), "obs2": torch.randn(B, O, generator=hg).pin_memory(),
                 "act": (torch.rand(B, A, generator=hg) * 0.8 - 0.4).pin_memory(), "rew": torch.randn(B, generator=hg).pin_memory(),
                 "done": torch.zeros(B).pin_memory()} for _ in range(4)]
        it = 0
        for i in range(20):
            alg.local_update(ring[i % 4], it); it += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prev = None
        sink = 0.0
        for i in range(a.steps):
            tb = alg.local_update(ring[i % 4], it); it += 1
            if read and prev is not None:
                sink += prev["Loss/Critic loss-RL iter"]
            prev = tb
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{tag}: {a.steps / dt:.0f} steps/s ({1e6 * dt / a.steps:.0f} us/step)")

    loop("default affinity, read every step")
    loop("default affinity, no stats read", read=False)
    torch.set_num_threads(4)
    loop("torch threads=4 (the reference's setting, utils/init_args.py:14)")
    torch.set_num_threads(1)
    loop("torch threads=1")
    gpu_node = sh(f"cat /sys/bus/pci/devices/{bdf}/numa_node")
    try:
        use = nodes[int(gpu_node)] & aff0
        if use:
            os.sched_setaffinity(0, use)
            loop(f"affinity = node {gpu_node}, fresh pinned buffers")
    except Exception as e:  # noqa: BLE001
        print("no node binding:", e)
    # host-side cost alone: device-resident batch (no H2D), same call path
    dev = {k: v.cuda() for k, v in {"obs": torch.randn(B, O), "obs2": torch.randn(B, O), "act": torch.rand(B, A) * 0.8 - 0.4,
                                    "rew": torch.randn(B), "done": torch.zeros(B)}.items()}
    it = 10_000
    for _ in range(20):
        alg.local_update(dev, it); it += 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        alg.local_update(dev, it); it += 1
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"device-resident batch through local_update: {a.steps / dt:.0f} steps/s; host issue time {1e6 * t_issue / a.steps:.0f} us/step")


if __name__ == "__main__":
    main()
