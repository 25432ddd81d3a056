"""NUMA placement for the host side of the host-buffer entry points (pplp_circuit_a_host, pplp_proximity_batch_host).

On a multi-socket GPU box the PCIe root of a GPU hangs off one socket; pinned buffers that live on the other socket make
every H2D/D2H copy cross the inter-socket fabric, and eight ranks sharing one socket's memory controllers cap the whole
box (round 1: end-to-end efficiency 0.30 at 8 GPUs).  bind_to_gpu_node() moves the calling process onto the CPUs of the
GPU's node (when the cpuset allows it) so that cudaHostAlloc — which places pages on the calling thread's node — and the
thread that feeds the copies are local.  Everything here is best effort and reports what it found."""
import ctypes
import os
import subprocess


def _read(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(index):
    """NUMA node of GPU `index` from sysfs (None when unknown, -1 when the platform reports none)."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception:
        return None
    if not bus:
        return None
    dom, rest = bus.split(":", 1)
    path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
    v = _read(path)
    return int(v) if v is not None and v.lstrip("-").isdigit() else None


def topology():
    nodes = {}
    base = "/sys/devices/system/node"
    if os.path.isdir(base):
        for d in sorted(os.listdir(base)):
            if d.startswith("node") and d[4:].isdigit():
                nodes[int(d[4:])] = sorted(_parse_cpulist(_read(os.path.join(base, d, "cpulist"))))
    status = _read("/proc/self/status") or ""
    mems = next((ln.split(":", 1)[1].strip() for ln in status.splitlines() if ln.startswith("Mems_allowed_list")), None)
    return {"nodes": {k: len(v) for k, v in nodes.items()}, "node_cpus": nodes, "mems_allowed": mems, "cpus_allowed": sorted(os.sched_getaffinity(0))}


def bind_to_gpu_node(index):
    """Pin this process to the allowed CPUs of GPU `index`'s NUMA node.  Returns a report dict (also when nothing was done)."""
    topo = topology()
    node = gpu_numa_node(index)
    rep = {"gpu": index, "gpu_numa_node": node, "numa_nodes": topo["nodes"], "mems_allowed": topo["mems_allowed"], "cpus_allowed": len(topo["cpus_allowed"]), "bound": False}
    if node is None or node < 0 or node not in topo["node_cpus"]:
        rep["why"] = "GPU NUMA node unknown or not exposed"
        return rep
    local = sorted(set(topo["node_cpus"][node]) & set(topo["cpus_allowed"]))
    if not local:
        rep["why"] = f"no allowed CPU on node {node} (cpuset restricts this container to other nodes)"
        return rep
    try:
        os.sched_setaffinity(0, local)
        rep["bound"] = True
        rep["cpus_bound"] = len(local)
    except OSError as e:
        rep["why"] = f"sched_setaffinity failed: {e}"
    return rep


def pinned_empty(lib, shape, dtype, write_combined=False):
    """A page-locked torch CPU tensor from cudaHostAlloc (pplp_host_alloc_ex), allocated on the calling thread's NUMA node."""
    import numpy as np
    import torch
    n = int(np.prod(shape))
    item = torch.empty((), dtype=dtype).element_size()
    ptr = ctypes.c_void_p()
    rc = lib.pplp_host_alloc_ex(n * item, 1 if write_combined else 0, ctypes.byref(ptr))
    if rc != 0:
        raise MemoryError("pplp_host_alloc_ex failed")
    buf = (ctypes.c_uint8 * (n * item)).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=dtype).reshape(shape)
    t._pplp_host_ptr = ptr        # keeps the address for pplp_host_free; the buffer lives until the process exits otherwise
    return t
