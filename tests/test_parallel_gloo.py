"""CPU / gloo, world_size 2 and 3: the sharding, broadcast and gather logic of
wisecondorx_b200.parallel.get_reference_sharded with the NumPy oracle injected as the compute step
(the product compute path is CUDA; this only exercises the N > 1 host logic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle
from wisecondorx_b200 import newref_tools, parallel, synth


def _oracle_compute(xd, per, cum, ref_size, start, end, sample_ids):
    x = xd.cpu().numpy()
    idx, dst = np_oracle.get_reference_topk(x, per, cum, ref_size, start, end)
    nr = np_oracle.null_ratios(x, idx, start, end, sample_ids)
    return torch.from_numpy(idx), torch.from_numpy(dst), torch.from_numpy(nr)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = [23, 19, 17, 13, 11] + [3] * 17
    x, per, cum = synth.make_corrected_matrix(per, 9, seed=4)
    ids = [3, 1, 8, 0]
    out = parallel.get_reference_sharded(x if rank == 0 else None, per, cum, 12, ids, device=torch.device("cpu"),
                                         compute_fn=_oracle_compute)
    if rank == 0:
        np.savez(tmp, idx=out[0], dist=out[1], nr=out[2])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_single(world, tmp_path):
    tmp = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), tmp), nprocs=world, join=True)
    got = np.load(tmp)
    per = [23, 19, 17, 13, 11] + [3] * 17
    x, per, cum = synth.make_corrected_matrix(per, 9, seed=4)
    idx, dst, nr = np_oracle.get_reference(x, per, cum, 12, 1, 1, [3, 1, 8, 0])
    assert np.array_equal(got["idx"], idx) and np.array_equal(got["dist"], dst)
    assert np.array_equal(got["nr"], nr, equal_nan=True)


def test_shard_bounds_are_the_reference_parts():
    for n, w in [(2815, 8), (191678, 8), (27941, 3)]:
        b = parallel.shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert b == [newref_tools._get_part(r, w, n) for r in range(w)]


def _worker_shm(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = [23, 19, 17, 13, 11] + [3] * 17
    x, per, cum = synth.make_corrected_matrix(per, 9, seed=4)
    ids = [3, 1, 8, 0]
    sr = parallel.ShardedReference(x.shape[0], x.shape[1], 12, len(ids), torch.device("cpu"), compute_fn=_oracle_compute)
    for _ in range(2):  # buffers are reused across runs
        out = sr.run(sr.slice_of(x), per, cum, ids)
    if rank == 0:
        np.savez(tmp, idx=out[0], dist=out[1], nr=out[2])
    sr.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reference_shared_segment(world, tmp_path):
    """ShardedReference: sliced upload + all-gather of X, per-rank parts, row blocks written by every rank into one
    shared host segment -- equals the single-process result."""
    tmp = str(tmp_path / "out.npz")
    mp.spawn(_worker_shm, args=(world, _free_port(), tmp), nprocs=world, join=True)
    got = np.load(tmp)
    per = [23, 19, 17, 13, 11] + [3] * 17
    x, per, cum = synth.make_corrected_matrix(per, 9, seed=4)
    idx, dst, nr = np_oracle.get_reference(x, per, cum, 12, 1, 1, [3, 1, 8, 0])
    assert np.array_equal(got["idx"], idx) and np.array_equal(got["dist"], dst)
    assert np.array_equal(got["nr"], nr, equal_nan=True)


def _worker_predict(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    samples = list(range(100, 111))  # 11 "samples": uneven blocks

    def fake(args, my, ref, gender, eng):
        return {"ids": list(my)}, [s % 7 for s in my]

    (a, b), res, summary = parallel.predict_batch_sharded(None, samples, None, "A", None, process_fn=fake)
    assert res["ids"] == samples[a:b]
    if rank == 0:
        np.save(tmp, np.array(summary))
    else:
        assert summary is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_predict_batch_sharded_over_samples(world, tmp_path):
    tmp = str(tmp_path / "sum.npy")
    mp.spawn(_worker_predict, args=(world, _free_port(), tmp), nprocs=world, join=True)
    assert np.load(tmp).tolist() == [s % 7 for s in range(100, 111)]
    blocks = parallel.shard_samples(11, world)
    assert blocks[0][0] == 0 and blocks[-1][1] == 11 and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
