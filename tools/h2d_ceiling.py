"""What the box allows: host-to-device bandwidth of N GPUs copying at the same time, against one GPU copying alone.

The end-to-end number of the benchmark (bench.py `e2e`) is bound by the 9.33 MB of wire data per frame crossing the
host-to-device link (leaf_grasp_node_v3.py:110-111 is the copy the path replaces), so its scaling over GPUs is the scaling
of this copy.  Every rank allocates the benchmark's input volume (256 frames: 0.80 GB of labels + 1.59 GB of depth) in
pinned host memory and copies it to its GPU with cudaMemcpyAsync:

  solo         one rank at a time, the others idle           -> what a GPU's link gives alone
  concurrent   all ranks together (barrier, then copy)       -> what the host side gives when every link is busy
  two_streams  as concurrent, labels and depth on two streams
  wc           as concurrent, from write-combined pinned memory (cudaHostAllocWriteCombined)

Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py
(or plainly `python tools/h2d_ceiling.py` for one GPU).  Rank 0 prints one JSON line.
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.distributed as dist

FRAMES, P = 256, 1440 * 1080
REPS = 4


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind = "--no-bind" not in sys.argv
    from leafgrasp_b200 import hostmem
    cpus = hostmem.bind_to_gpu(local) if bind else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = C.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    n_lab, n_dep = FRAMES * P * 2, FRAMES * P * 4
    dev = torch.empty(n_lab + n_dep, dtype=torch.uint8, device="cuda")

    def host_buffer(flags):
        p = C.c_void_p()
        rc = rt.cudaHostAlloc(C.byref(p), n_lab + n_dep, flags)
        if rc != 0:
            raise RuntimeError(f"cudaHostAlloc({flags}) -> {rc}")
        C.memset(p, 1, n_lab + n_dep)          # first touch by this (bound) process
        return p

    plain = host_buffer(0)
    wc = host_buffer(4)                        # cudaHostAllocWriteCombined
    s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def copy_once(src, two):
        a, b = (s0, s1) if two else (s0, s0)
        rt.cudaMemcpyAsync(dev.data_ptr(), src, n_lab, 1, C.c_void_p(a.cuda_stream))
        rt.cudaMemcpyAsync(dev.data_ptr() + n_lab, C.c_void_p(src.value + n_lab), n_dep, 1, C.c_void_p(b.cuda_stream))

    def timed(src, two):
        copy_once(src, two)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(REPS):
            copy_once(src, two)
        s0.synchronize(); s1.synchronize()
        dt = time.perf_counter() - t0
        return REPS * (n_lab + n_dep) / dt / 1e9

    def gather(v):
        if world == 1:
            return [v]
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    solo = 0.0
    for r in range(world):                     # one rank at a time
        barrier()
        if r == rank:
            copy_once(plain, False); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(REPS):
                copy_once(plain, False)
            s0.synchronize()
            solo = REPS * (n_lab + n_dep) / (time.perf_counter() - t0) / 1e9
        barrier()
    res = {"solo": gather(solo)}
    for name, src, two in (("concurrent", plain, False), ("two_streams", plain, True), ("wc", wc, False)):
        res[name] = gather(timed(src, two))
        barrier()
    if rank == 0:
        line = {"tool": "h2d_ceiling", "n_gpus": world, "bytes_per_copy": n_lab + n_dep, "reps": REPS,
                "bound_to_gpu_cpus": bool(cpus), "cpus_rank0": len(cpus) if cpus else None,
                "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])}
        for k, v in res.items():
            line[k + "_gbs_per_gpu"] = [round(x, 1) for x in v]
            line[k + "_gbs_total"] = round(sum(v), 1)
        line["frames_per_s_at_concurrent_ceiling"] = round(line["concurrent_gbs_total"] * 1e9 / (P * 6), 0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
