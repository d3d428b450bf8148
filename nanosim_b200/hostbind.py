"""Host-side placement of a GPU process: run its threads, and with them the pinned buffers they allocate, on the CPUs of
the NUMA node the GPU's PCIe link hangs off.  Device->host copies into pinned memory of the OTHER socket cross the
inter-socket link at a fraction of the PCIe rate, which is what limits the end-to-end rate of 4-8 ranks on a two-socket
box (one rank per GPU, every rank streaming GBs per second into host memory).

No counterpart in the reference (its workers only use CPUs); the multi-process fan-out it replaces is
/root/reference/src/simulator.py:1588-1639.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_address(device):
    """'0000:1b:00.0'-style sysfs name of CUDA device `device` (None when it cannot be found)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        if hasattr(p, "pci_bus_id"):
            return "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    except Exception:
        pass
    try:
        import subprocess
        out = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        if out:                                   # 00000000:1B:00.0 -> 0000:1b:00.0
            dom, rest = out.split(":", 1)
            return (dom[-4:] + ":" + rest).lower()
    except Exception:
        pass
    return None


def gpu_local_cpus(device):
    """(numa node, set of CPUs local to the GPU) from sysfs; (None, None) when unknown or the box has one node."""
    addr = gpu_pci_address(device)
    if not addr:
        return None, None
    base = "/sys/bus/pci/devices/" + addr
    try:
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
    except (OSError, ValueError):
        return None, None
    if node < 0 or not cpus:
        return None, None
    return node, cpus


def bind_to_gpu_node(device, verbose=False):
    """Restricts the calling thread (and every thread it starts later) to the CPUs of the GPU's NUMA node.  Returns a dict
    describing what was done; a no-op ({"bound": False, ...}) on single-node hosts, without sysfs, or with
    NANOSIM_B200_NO_BIND=1."""
    info = {"bound": False, "node": None, "cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}
    if os.environ.get("NANOSIM_B200_NO_BIND") or not hasattr(os, "sched_setaffinity"):
        return info
    node, cpus = gpu_local_cpus(device)
    if node is None:
        return info
    allowed = os.sched_getaffinity(0)
    target = cpus & allowed
    if not target or target == allowed:
        info["node"] = node
        return info
    try:
        os.sched_setaffinity(0, target)
    except OSError:
        return info
    info.update(bound=True, node=node, cpus=len(target))
    if verbose:
        print("nanosim_b200: GPU %d is on NUMA node %d: process bound to its %d CPUs" % (device, node, len(target)))
    return info


def unbind(all_cpus=None):
    """Back to every CPU of the machine (e.g. before forking CPU-only workers)."""
    if not hasattr(os, "sched_setaffinity"):
        return
    try:
        os.sched_setaffinity(0, all_cpus or range(os.cpu_count() or 1))
    except OSError:
        pass
