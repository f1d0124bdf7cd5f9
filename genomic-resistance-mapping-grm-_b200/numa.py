"""Host-memory placement for the PCIe legs of the path (H2D of the text, D2H of the matrix).

Page-locked buffers land on the NUMA node of the thread that first touches them.  On a multi-socket box a rank
whose staging memory sits on the other socket pays a cross-socket hop for every byte it sends to its GPU, and
eight ranks doing that at once saturate the inter-socket link (round 1: e2e efficiency 0.31 at N = 8).
``near_gpu(device)`` is a context manager that pins the calling thread to the CPUs of the GPU's own NUMA node
while the buffers are allocated and touched, then restores the previous affinity.
"""
from __future__ import annotations

import contextlib
import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(device: int) -> int | None:
    """NUMA node of CUDA device `device` from sysfs, or None when the platform does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node: int) -> set[int]:
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            return _parse_cpulist(f.read())
    except Exception:
        return set()


@contextlib.contextmanager
def near_gpu(device: int):
    """Run the body on the CPUs of the GPU's NUMA node (first-touch places what it allocates there)."""
    old = None
    try:
        node = gpu_numa_node(device)
        cpus = node_cpus(node) if node is not None else set()
        if cpus and hasattr(os, "sched_getaffinity"):
            old = os.sched_getaffinity(0)
            use = cpus & old or cpus
            os.sched_setaffinity(0, use)
        yield node
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except OSError:
                pass
