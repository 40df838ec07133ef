"""Host-side placement helper for the end-to-end (host buffer) path: pinned staging buffers should live on the NUMA
node the GPU hangs off, otherwise concurrent device-to-host copies of several ranks share the inter-socket link."""
from __future__ import annotations

import os


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device) -> int | None:
    """NUMA node of a CUDA device from sysfs (/sys/bus/pci/devices/<bus id>/numa_node), None when unknown."""
    import torch

    try:
        p = torch.cuda.get_device_properties(device)
        bus_id = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus_id}/numa_node").read())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001  (no sysfs entry, virtualised PCI topology, ...)
        return None


def bind_to_gpu_numa_node(device) -> dict:
    """Restrict the calling process to the CPUs of `device`'s NUMA node (intersected with what it may already use),
    so that the pinned buffers it allocates afterwards are first-touched, hence placed, on that node. Call before the
    first host-buffer step. Returns what was done (for the bench line)."""
    node = gpu_numa_node(device)
    if node is None:
        return {"node": None, "bound": False}
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            return {"node": node, "bound": False}
        os.sched_setaffinity(0, target)
        return {"node": node, "bound": True, "cpus": len(target)}
    except Exception:  # noqa: BLE001
        return {"node": node, "bound": False}
