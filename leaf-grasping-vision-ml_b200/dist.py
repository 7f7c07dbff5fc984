"""Multi-GPU plumbing: frames are independent, so a batch is split into contiguous
per-rank ranges, every rank runs the whole path on its range, and the only exchange is one all-gather of
fixed-size candidate records for the aggregate result.  torch.distributed (NCCL on GPUs, gloo in the CPU
tests) carries it; there is no data-path collective to fuse with."""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_FIELDS = 4   # x, y, traditional score, ml score  (320 B per frame at 20 candidates)


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous block [lo, hi) of frame indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def records_from_results(res, device) -> torch.Tensor:
    """Structured FRAME_RESULT array -> float32 tensor [frames, 20, 4] (x, y, trad, ml)."""
    import numpy as np
    rec = np.stack([res["cand_x"].astype(np.float32), res["cand_y"].astype(np.float32),
                    res["trad"].astype(np.float32), np.nan_to_num(res["ml"]).astype(np.float32)], axis=-1)
    unused = np.arange(rec.shape[1])[None, :] >= res["n_candidates"][:, None]
    rec[unused] = (-1.0, -1.0, 0.0, 0.0)        # same convention as the records the fusion kernel writes
    return torch.from_numpy(rec).to(device)


def records_from_result_buffer(buf: torch.Tensor, n_frames: int) -> torch.Tensor:
    """Same records, built from the raw lg_frame_result byte buffer WITHOUT leaving the device (no host sync):
    buf is the uint8 tensor lg_process_batch filled (GraspEngine.process_batch(..., sync=False))."""
    from . import _native as N
    dt = N.FRAME_RESULT
    rows = buf.reshape(n_frames, dt.itemsize)

    def field(name, tdtype):
        off = dt.fields[name][1]
        nbytes = dt.fields[name][0].itemsize
        return rows[:, off:off + nbytes].contiguous().view(tdtype)

    x = field("cand_x", torch.int32).to(torch.float32)
    y = field("cand_y", torch.int32).to(torch.float32)
    trad = field("trad", torch.float64).to(torch.float32)
    ml = torch.nan_to_num(field("ml", torch.float64)).to(torch.float32)
    rec = torch.stack([x, y, trad, ml], dim=-1)
    ncand = field("n_candidates", torch.int32)
    unused = torch.arange(rec.shape[1], device=rec.device)[None, :] >= ncand
    rec[unused] = torch.tensor([-1.0, -1.0, 0.0, 0.0], device=rec.device)
    return rec


def gather_records_async(local: torch.Tensor, out: "torch.Tensor | None" = None):
    """ONE all-gather of equally sized per-rank record blocks [frames_per_rank, 20, 4] (the buffer the fusion kernel
    filled through GraspEngine.set_record_output) into out [world * frames_per_rank, 20, 4], issued asynchronously:
    NCCL runs it on its own stream after the work already queued on the current stream, so it overlaps whatever the
    caller launches next.  Returns (out, work); work.wait() orders the current stream after the gather (work is None
    without a process group)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local, None
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(out, local, async_op=True)
    return out, work


def gather_candidate_records(local: torch.Tensor, n_frames: int) -> torch.Tensor:
    """All-gather of the per-rank record blocks into [n_frames, 20, 4] on every rank (blocks may differ in size by one
    frame: shard_range)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_frames + world - 1) // world
    pad = torch.zeros(per, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * per, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_frames, r, world)
        parts.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(parts, dim=0)
