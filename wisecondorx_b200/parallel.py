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
    engine.reference(start, end, ref_size, sample_ids, device_out=(idx.data_ptr(), dst.data_ptr(), nr.data_ptr()))
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


class ShardedReference:
    """Host-to-host get_reference over the ranks of one box with every PCIe link and NVLink used once:

      1. every rank copies ITS slice of X (ceil(N / W) rows) from its host memory to its GPU  -> H2D time / W
      2. one NCCL all-gather of the slices over NVLink / NVSwitch gives every GPU the whole matrix (the only
         exchange of the data path: each target bin needs all of X but no other bin's result)
      3. every rank runs its part (the reference's part r + 1 of W, newref_tools.py:244-247)
      4. every rank copies its row block straight into ONE host result array that lives in a POSIX
         shared-memory segment mapped (and page-locked) by all ranks                       -> D2H time / W
      5. barrier: rank 0 holds the complete (indexes, distances, null_ratios) in part order

    The segment, the registrations and the device buffers are created once and reused by every run() (what a
    pinned staging buffer is at N = 1).  `compute_fn` as in get_reference_sharded (CPU / gloo tests)."""

    def __init__(self, n, s, ref_size, m, device, group=None, compute_fn=None, engine=None):
        import os
        from multiprocessing import shared_memory
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n, self.s, self.k, self.m = int(n), int(s), int(ref_size), int(m)
        self.device = torch.device(device)
        self.compute_fn = compute_fn
        self.bounds = shard_bounds(self.n, self.world)
        self.slice_rows = -(-self.n // self.world)
        self.x_full = torch.empty((self.slice_rows * self.world, self.s), dtype=torch.float64, device=self.device)
        self.x_slice = torch.empty((self.slice_rows, self.s), dtype=torch.float64, device=self.device)
        start, end = self.bounds[self.rank]
        rows = end - start
        self.dev_out = (torch.empty((rows, self.k), dtype=torch.int32, device=self.device),
                        torch.empty((rows, self.k), dtype=torch.float64, device=self.device),
                        torch.empty((rows, self.m), dtype=torch.float64, device=self.device))
        sizes = [self.n * self.k * 4, self.n * self.k * 8, self.n * self.m * 8]
        offs = [0, sizes[0], sizes[0] + sizes[1]]
        total = max(1, sum(sizes))
        name = [None]
        if self.rank == 0:
            try:
                self._shm = shared_memory.SharedMemory(create=True, size=total)
                os.posix_fallocate(self._shm._fd, 0, total)  # a too small /dev/shm fails here, not with SIGBUS later
                name[0] = self._shm.name
            except OSError as e:
                name[0] = None
                self._err = str(e)
        dist.broadcast_object_list(name, 0, group=group)
        if name[0] is None:  # every rank raises: callers fall back to get_reference_sharded (gather through rank 0)
            raise RuntimeError("ShardedReference: cannot create the shared result segment ({} bytes)".format(total))
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
        buf = self._shm.buf
        self.host_out = (np.ndarray((self.n, self.k), dtype=np.int32, buffer=buf, offset=offs[0]),
                         np.ndarray((self.n, self.k), dtype=np.float64, buffer=buf, offset=offs[1]),
                         np.ndarray((self.n, self.m), dtype=np.float64, buffer=buf, offset=offs[2]))
        self._registered = False
        if self.device.type == "cuda":
            ptr = self.host_out[0].ctypes.data
            rc = torch.cuda.cudart().cudaHostRegister(ptr, total, 0)
            self._registered = int(rc) == 0
            self._reg_ptr = ptr
            if compute_fn is None:
                self.engine = engine or newref_tools.NewrefEngine(self.device.index or 0)
                self.engine.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._host_t = tuple(torch.from_numpy(a) for a in self.host_out)
        self.chunks = 4 if self.device.type == "cuda" and self.world > 1 else 1
        if self.chunks > 1:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._copy_ev = [torch.cuda.Event() for _ in range(self.chunks)]
        dist.barrier(group=group)

    def slice_of(self, x):
        """This rank's slice of the host matrix (rows [rank * slice_rows, ...)), e.g. to pin it once."""
        a = self.rank * self.slice_rows
        return x[a: min(self.n, a + self.slice_rows)]

    def run(self, x_slice_host, per, cum, sample_ids, phases=None):
        """x_slice_host: this rank's slice of X (torch CPU tensor, ideally pinned, or NumPy).  Returns the three host
        arrays (views of the shared segment; complete on every rank after the closing barrier).
        phases: optional dict that receives the wall-clock of the phases in ms (adds stream synchronisations)."""
        import time
        t_start = time.perf_counter()

        def mark(name):
            if phases is not None:
                if self.device.type == "cuda":
                    torch.cuda.synchronize(self.device)
                phases[name] = (time.perf_counter() - t_start) * 1e3
        if not torch.is_tensor(x_slice_host):
            x_slice_host = torch.from_numpy(np.ascontiguousarray(x_slice_host, dtype=np.float64))
        r = x_slice_host.shape[0]
        if self.world > 1 and self.device.type == "cuda" and self.chunks > 1:
            # upload and exchange in row chunks: the NCCL all-gather of chunk c runs (on NCCL's stream) while chunk
            # c + 1 is still crossing PCIe; every chunk lands at its final place in the gathered matrix
            cur = torch.cuda.current_stream(self.device)
            step = -(-self.slice_rows // self.chunks)
            for c in range(self.chunks):
                a, b = c * step, min(self.slice_rows, (c + 1) * step)
                if b <= a:
                    break
                with torch.cuda.stream(self._copy_stream):
                    if a < r:
                        self.x_slice[a:min(b, r)].copy_(x_slice_host[a:min(b, r)], non_blocking=True)
                    self._copy_ev[c].record(self._copy_stream)
                cur.wait_event(self._copy_ev[c])
                outs = [self.x_full[w * self.slice_rows + a: w * self.slice_rows + b] for w in range(self.world)]
                dist.all_gather(outs, self.x_slice[a:b], group=self.group)
        else:
            self.x_slice[:r].copy_(x_slice_host, non_blocking=True)
            if self.world > 1:
                dist.all_gather_into_tensor(self.x_full, self.x_slice, group=self.group)
            else:
                self.x_full.copy_(self.x_slice)
        mark("upload_and_all_gather")
        xd = self.x_full[: self.n]
        start, end = self.bounds[self.rank]
        ids = list(sample_ids)
        if self.compute_fn is None:
            self.engine.load(None, per, cum, on_device_ptr=xd.data_ptr(), shape=(self.n, self.s))
            mark("prepare_operands")
            # host outputs: the library copies every finished row block to the (page-locked) shared segment while the
            # next block is still in the re-rank, instead of one D2H of the whole part at the end
            self.engine.reference(start, end, self.k, ids, out=tuple(a[start:end] for a in self.host_out))
        else:
            outs = self.compute_fn(xd, np.asarray(per, dtype=np.int64), np.asarray(cum, dtype=np.int64), self.k, start, end, ids)
            for host, dev in zip(self._host_t, outs):
                host[start:end].copy_(dev, non_blocking=True)
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()
        mark("compute_and_download")
        dist.barrier(group=self.group)
        mark("barrier")
        return self.host_out

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self._reg_ptr)
            self._registered = False
        self._host_t = None
        self.host_out = None
        dist.barrier(group=self.group)
        try:
            self._shm.close()
        except BufferError:  # a caller still holds a view of the result arrays: the mapping goes with its last reference
            pass
        if self.rank == 0:
            try:
                self._shm.unlink()
            except FileNotFoundError:
                pass


# ---------------------------------------------------------------------------------------------
# predict, batch of samples (BASELINE config 5): samples sharded over the ranks, reference replicated
# ---------------------------------------------------------------------------------------------
def shard_samples(n_samples: int, world: int):
    """Contiguous blocks of samples per rank (same arithmetic as the target-bin parts)."""
    return [newref_tools._get_part(r, world, n_samples) for r in range(world)]


def predict_batch_sharded(args, samples, ref_file, ref_gender="A", engine=None, process_fn=None, group=None):
    """`normalize` + CBS for a batch of samples with the SAMPLES sharded over the ranks.  Every rank holds the
    whole reference (0.7 GB at 15 kb) and processes its block of samples independently -- no collective on the data
    path, exactly like the reference's one-process-per-sample shell loop (docs/include/pipeline/predict.sh:15-23);
    each rank keeps (and would write out) the results of its own samples.  Rank 0 additionally receives a small
    summary (segments per sample) through gather_object.

    process_fn(args, my_samples, ref_file, ref_gender, engine) -> (results, segments_per_sample); defaults to
    normalize_batch + segment_batch on the GPU (the gloo tests inject a CPU function)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    a, b = shard_samples(len(samples), world)[rank]
    mine = samples[a:b]
    if process_fn is None:
        from . import predict_control

        def process_fn(args_, my, ref, gender, eng):
            if not my:
                return None, []
            r, z, w, nref, m_lr, m_z = predict_control.normalize_batch(args_, my, ref, gender, eng)
            offs = np.concatenate([[0], np.asarray(ref["masked_bins_per_chr_cum"], dtype=np.int64)])
            ends = predict_control.segment_batch(r, w, nref, m_lr, offs, getattr(args_, "minrefbins", 150), getattr(args_, "alpha", 1e-4),
                                                 10000, getattr(args_, "seed", None) or 0, eng.ctx if eng else None)
            nchr = len(offs) - 1
            per_sample = [int(sum(len(e) for e in ends[i * nchr:(i + 1) * nchr])) for i in range(len(my))]
            return {"r": r, "z": z, "w": w, "ref_sizes": nref, "m_lr": m_lr, "m_z": m_z, "segment_ends": ends}, per_sample

    results, per_sample = process_fn(args, mine, ref_file, ref_gender, engine)
    summary = None
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((a, b, per_sample), gathered, dst=0, group=group)
        if rank == 0:
            summary = [n for (_, _, lst) in sorted(gathered, key=lambda t: t[0]) for n in lst]
    else:
        summary = list(per_sample)
    return (a, b), results, summary
