"""Host-side NUMA placement for the host -> device leg of the update path.

A 12.6 MB minibatch per step (Humanoid, B = 4096) crosses PCIe from pinned host memory.  If the process (and therefore
its first-touched pinned pages) sits on the other socket than the GPU, the DMA crosses the inter-socket link and the
end-to-end rate halves (measured on the 2-socket B200 hosts of this pool: 27 GB/s instead of 49 GB/s).  `bind_to_gpu_node`
restricts the calling process to the CPUs of the GPU's NUMA node; call it BEFORE allocating pinned buffers.  Opt-in:
`bench.py` calls it for its own process; the drop-in does it only when asked (`dsact_numa_bind=True`).
"""
from __future__ import annotations

import glob
import os
from typing import Optional


def _node_cpus(node: int):
    cpus = set()
    try:
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
    except OSError:
        pass
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of CUDA device `device_index` (sysfs, via its PCI address), or None if it cannot be told."""
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:   # noqa: BLE001  (no sysfs, virtualised PCI topology, old torch ...)
        return None


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns what was done (for logs / bench lines)."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    if not hasattr(os, "sched_setaffinity") or not glob.glob("/sys/devices/system/node/node[0-9]*"):
        return info
    node = gpu_numa_node(device_index)
    info["numa_node"] = node
    if node is None:
        return info
    use = _node_cpus(node) & os.sched_getaffinity(0)
    if use:
        os.sched_setaffinity(0, use)
        info["cpus"], info["bound"] = len(use), True
    return info
