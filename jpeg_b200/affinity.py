"""CPU / memory affinity for one-process-per-GPU jobs.  The end-to-end path moves 3 bytes per pixel through pinned host memory;
on a two-socket host a rank whose staging buffers live on the other socket pays the inter-socket link for every byte (measured
in round 1: 39 GB/s per GPU at N = 4 against 51 GB/s alone).  Binding a rank's threads to the NUMA node of its GPU before the
pinned buffers are allocated keeps them local (first-touch placement).  Everything here is best effort: without the sysfs
entries, or when the node's CPUs are outside the process's cpuset, nothing changes."""
from __future__ import annotations

import os


def parse_cpulist(text: str):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of /sys/devices/system/node/node*/cpulist)"""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_address(domain: int, bus: int, device: int) -> str:
    return f"{domain:04x}:{bus:02x}:{device:02x}.0"


def numa_node_of(pci: str, sysfs: str = "/sys"):
    """NUMA node of a PCI device, or None when the kernel does not say (file missing or -1)"""
    try:
        with open(os.path.join(sysfs, "bus", "pci", "devices", pci, "numa_node")) as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def cpus_of_node(node: int, sysfs: str = "/sys"):
    try:
        with open(os.path.join(sysfs, "devices", "system", "node", f"node{node}", "cpulist")) as f:
            return parse_cpulist(f.read())
    except (OSError, ValueError):
        return set()


def plan(pci: str, allowed, sysfs: str = "/sys"):
    """-> (node, cpus to bind to) or (None, None) when there is nothing sensible to do"""
    node = numa_node_of(pci, sysfs)
    if node is None:
        return None, None
    cpus = cpus_of_node(node, sysfs) & set(allowed)
    if not cpus or cpus == set(allowed):
        return node, None
    return node, cpus


def bind_to_gpu(device_index: int, sysfs: str = "/sys") -> str:
    """Binds the calling process to the CPUs of the GPU's NUMA node; returns a one-line description for the bench record."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        pci = pci_address(p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        allowed = os.sched_getaffinity(0)
        node, cpus = plan(pci, allowed, sysfs)
        if node is None:
            return f"unchanged (no NUMA information for {pci})"
        if cpus is None:
            return f"unchanged (GPU on node {node}; its CPUs are all or none of this process's cpuset)"
        os.sched_setaffinity(0, cpus)
        return f"node {node}: {len(cpus)} of {len(allowed)} CPUs"
    except Exception as e:  # never let a placement hint break a run
        return f"unchanged ({type(e).__name__}: {e})"
