"""Multi-GPU newref: the target-bin axis of get_reference sharded over the ranks of a
torch.distributed process group (one process per GPU, NCCL over NVLink).

The reference's only parallel axis is the same one: `get_reference(part, split_parts)` handles the
bins `_get_part(part - 1, split_parts, N)` (newref_tools.py:168, :244-247) and the parts are
concatenated in part order (newref_control.py:165-174).  Here rank r of W takes part r + 1 of W, so
the gathered result is identical to a single-GPU run with cpus = 1 given the same null-sample draw.

Data path: every target bin needs the whole matrix X but no other bin's result, so the only exchange
is X itself (rank 0 -> all: one NCCL broadcast, 0.77 GB at 15 kb / 500 samples) and the gather of the
row blocks to rank 0.  There is no collective inside the distance sweep.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import newref_tools


def shard_bounds(n: int, world: int):
    """Row range of every rank = the reference's parts (newref_tools.py:244-247)."""
    return [newref_tools._get_part(r, world, n) for r in range(world)]


def _gpu_compute(x_dev, per, cum, ref_size, start, end, sample_ids, engine):
    n, s = x_dev.shape
    engine.load(None, per, cum, on_device_ptr=x_dev.data_ptr(), shape=(n, s))
    rows, m = end - start, len(sample_ids)
    idx = torch.empty((rows, ref_size), dtype=torch.int32, device=x_dev.device)
    dst = torch.empty((rows, ref_size), dtype=torch.float64, device=x_dev.device)
    nr = torch.empty((rows, m), dtype=torch.float64, device=x_dev.device)
    engine.topk(start, end, ref_size, device_out=(idx.data_ptr(), dst.data_ptr()))
    engine.null_ratios(start, end, ref_size, sample_ids, device_out=nr.data_ptr())
    return idx, dst, nr


def get_reference_sharded(x, per, cum, ref_size, sample_ids, device=None, compute_fn=None, engine=None, group=None):
    """Sharded get_reference.  `x`: float64 [N, S] NumPy array on rank 0 (other ranks may pass None
    together with its shape through `per`/`cum` only -- the shape is broadcast).  Returns
    (indexes, distances, null_ratios) as NumPy arrays on rank 0 and None elsewhere.

    `compute_fn(x_tensor, per, cum, ref_size, start, end, sample_ids) -> (idx, dist, nr)` tensors on
    the tensor's device; defaults to the CUDA engine.  (The CPU/gloo tests inject the oracle here to
    exercise the sharding, broadcast and gather logic without a GPU.)"""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    cum = np.asarray(cum, dtype=np.int64)
    per = np.asarray(per, dtype=np.int64)
    shape = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == 0:
        shape[0], shape[1] = x.shape[0], x.shape[1]
    if world > 1:
        dist.broadcast(shape, 0, group=group)
    n, s = int(shape[0]), int(shape[1])
    xd = torch.empty((n, s), dtype=torch.float64, device=device)
    if rank == 0:
        xd.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)), non_blocking=True)
    if world > 1:
        dist.broadcast(xd, 0, group=group)
    bounds = shard_bounds(n, world)
    start, end = bounds[rank]
    if compute_fn is None:
        engine = engine or newref_tools.NewrefEngine(device.index or 0)
        if device.type == "cuda":
            engine.ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
        idx, dst, nr = _gpu_compute(xd, per, cum, ref_size, start, end, list(sample_ids), engine)
    else:
        idx, dst, nr = compute_fn(xd, per, cum, ref_size, start, end, list(sample_ids))
    if world == 1:
        return idx.cpu().numpy(), dst.cpu().numpy(), nr.cpu().numpy()
    return gather_row_blocks((idx, dst, nr), bounds, rank, device, group)


def gather_row_blocks(tensors, bounds, rank, device, group=None):
    """Gathers ragged row blocks to rank 0 in rank order (= the reference's part concatenation,
    newref_control.py:165-174).  Collectives need equal shapes: blocks are padded to the longest part."""
    max_rows = max(b[1] - b[0] for b in bounds)
    outs = []
    for t in tensors:
        pad = torch.zeros((max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
        pad[: t.shape[0]] = t
        if rank == 0:
            bufs = [torch.empty_like(pad) for _ in bounds]
            dist.gather(pad, bufs, 0, group=group)
            outs.append(torch.cat([buf[: b[1] - b[0]] for buf, b in zip(bufs, bounds)]).cpu().numpy())
        else:
            dist.gather(pad, None, 0, group=group)
    return tuple(outs) if rank == 0 else None
