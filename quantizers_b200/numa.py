"""Host-side NUMA placement for the staging buffers of one GPU.

A B200 box has its eight GPUs behind two CPU sockets; pinned staging memory that lands on the far socket makes every H2D / D2H
copy cross the socket interconnect, which eight concurrent ranks saturate long before their PCIe links.  Linux places pages on
the node of the CPU that first touches them, so it is enough to restrict the calling thread to the CPUs next to the GPU *before*
it allocates pinned memory (``cudaHostAlloc`` populates the pages in the calling thread).  No libnuma: the CPU list comes from
sysfs (``/sys/bus/pci/devices/<bdf>/local_cpulist``).

The reference has no counterpart (single process, one ``cuda:0``) -- this belongs to the one-process-per-GPU
layout of this package.
"""
from __future__ import annotations

import contextlib
import os
from typing import Iterator, Optional, Set


def parse_cpulist(text: str) -> Set[int]:
    """``"0-3,8,10-11"`` -> {0, 1, 2, 3, 8, 10, 11}; empty / malformed input -> empty set."""
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        try:
            if "-" in part:
                a, b = part.split("-", 1)
                cpus.update(range(int(a), int(b) + 1))
            else:
                cpus.add(int(part))
        except ValueError:
            return set()
    return cpus


def device_bdf(device_index: int) -> Optional[str]:
    import torch

    try:
        p = torch.cuda.get_device_properties(device_index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def local_cpus(device_index: int, sysfs: str = "/sys/bus/pci/devices") -> Set[int]:
    """CPUs on the NUMA node the GPU hangs off, intersected with what this process may run on; empty when unknown."""
    bdf = device_bdf(device_index)
    if bdf is None:
        return set()
    try:
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as f:
            cpus = parse_cpulist(f.read())
    except OSError:
        return set()
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return set()
    cpus &= allowed
    # a node-less box reports every CPU: nothing to gain, leave the thread alone
    return set() if cpus == allowed else cpus


@contextlib.contextmanager
def near_device(device_index: int) -> Iterator[bool]:
    """Run the body with the calling thread restricted to the GPU's local CPUs (allocate pinned buffers inside); the previous
    affinity is restored afterwards.  Yields whether a restriction was applied."""
    cpus = set() if os.environ.get("B200Q_NO_NUMA") == "1" else local_cpus(device_index)
    if not cpus:
        yield False
        return
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        yield False
        return
    try:
        yield True
    finally:
        try:
            os.sched_setaffinity(0, old)
        except OSError:
            pass
