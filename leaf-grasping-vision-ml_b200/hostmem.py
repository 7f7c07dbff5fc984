"""Host-side staging for the host-buffer entry point (``lg_process_batch_host``).

The wire data of the path (msg/masks.msg, msg/depth.msg -> leaf_grasp_node_v3.py:185-205) is 9.33 MB per
1440x1080 frame, so end to end the path is bound by the host-to-device link.  On a two-socket box a GPU
reaches its full link rate only from pinned memory of the socket it hangs off: one process per GPU therefore
binds itself to the CPUs NVML reports as local to its GPU *before* it allocates its pinned staging buffers
(first touch places the pages on that NUMA node).  Nothing here computes anything on the path.
"""
from __future__ import annotations

import os

import torch


def gpu_cpu_affinity(device_index: int):
    """CPUs local to the GPU (NVML's ideal affinity), or None when NVML cannot say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            # torch device indices follow CUDA_VISIBLE_DEVICES; NVML's do not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if device_index < len(ids) and ids[device_index].isdigit():
                    phys = int(ids[device_index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            return sorted(cpus) or None
        finally:
            pynvml.nvmlShutdown()
    except Exception:  # noqa: BLE001 - no NVML, no permission, single-socket container ...
        return None


def bind_to_gpu(device_index: int):
    """Pin the calling process to the CPUs next to its GPU.  Returns the CPU list used (None = left unbound)."""
    cpus = gpu_cpu_affinity(device_index)
    if not cpus:
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return cpus


def numa_node_of_cpus(cpus):
    """NUMA node(s) the given CPUs belong to, from sysfs (for the benchmark's report)."""
    nodes = set()
    try:
        for node in os.listdir("/sys/devices/system/node"):
            if not node.startswith("node"):
                continue
            with open(f"/sys/devices/system/node/{node}/cpulist") as fh:
                for part in fh.read().strip().split(","):
                    if not part:
                        continue
                    lo, _, hi = part.partition("-")
                    rng = range(int(lo), int(hi or lo) + 1)
                    if any(c in rng for c in cpus):
                        nodes.add(int(node[4:]))
    except OSError:
        pass
    return sorted(nodes)


def pinned_like(array) -> torch.Tensor:
    """A pinned host tensor holding `array` (NumPy), allocated and first touched by the calling (bound) process."""
    t = torch.from_numpy(array)
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t)
    return out
