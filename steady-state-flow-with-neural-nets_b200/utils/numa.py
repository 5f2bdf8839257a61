"""Host-side placement for the serving / training loops: pin the calling process to the CPU cores of the NUMA node its GPU
hangs off BEFORE it allocates pinned host buffers, so that the per-step uploads and downloads (2 MB in, 16.8 MB out per
64-image batch) cross one PCIe root complex and land in local DRAM.  With 8 ranks on a two-socket host the downloads of
round 1 reached 13 GB/s per GPU because half of them wrote across the socket interconnect (SCALE_r01: e2e efficiency 0.45).
Nothing here is on the GPU path; every failure degrades to "not pinned"."""
import os
import subprocess


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = []
    for part in s.split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_bus_id(device_index):
    """'0000:1b:00.0' of CUDA device `device_index` of this process (honours CUDA_VISIBLE_DEVICES), or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        dom, bus, dev = getattr(p, "pci_domain_id", None), getattr(p, "pci_bus_id", None), getattr(p, "pci_device_id", None)
        if bus is not None and dev is not None:
            return "%04x:%02x:%02x.0" % (dom or 0, bus, dev)
    except Exception:
        pass
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = vis.split(",")[device_index] if vis else str(device_index)
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", phys],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
        if out:
            bdf = out[0].strip().lower()
            return bdf[-12:] if len(bdf) > 12 else bdf          # nvidia-smi prints an 8-digit domain
    except Exception:
        pass
    return None


def gpu_numa_node(device_index):
    bdf = gpu_pci_bus_id(device_index)
    if bdf is None:
        return None
    node = _read("/sys/bus/pci/devices/%s/numa_node" % bdf)
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa_node(device_index, ranks_per_node_hint=None, local_rank=0):
    """Restrict this process to the cores of the GPU's NUMA node (a disjoint slice of them per local rank when
    ranks_per_node_hint is given).  Returns a dict describing what was done (for the bench line) or None."""
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    cpulist = _read("/sys/devices/system/node/node%d/cpulist" % node)
    if not cpulist:
        return None
    cpus = _parse_cpulist(cpulist)
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
    except (AttributeError, OSError):
        return None
    return {"numa_node": node, "cpus": len(allowed)}
