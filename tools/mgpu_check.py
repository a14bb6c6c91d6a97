"""torchrun check: get_reference sharded over the ranks (NCCL) == single-GPU result on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisecondorx_b200 import newref_tools, parallel, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = synth.config_bins(2)
x, per, cum = synth.make_corrected_matrix(per, 100, seed=2)
ids = list(range(0, 100, 3))
out = parallel.get_reference_sharded(x if rank == 0 else None, per, cum, 300, ids, device=dev)
if rank == 0:
    eng = newref_tools.NewrefEngine(local)
    eng.load(x, per, cum)
    n = x.shape[0]
    idx, dst = eng.topk(0, n, 300)
    nr = eng.null_ratios(0, n, 300, ids)
    ok = np.array_equal(out[0], idx) and np.array_equal(out[1], dst) and np.allclose(out[2], nr, rtol=1e-13, equal_nan=True)
    print("MGPU_CHECK", "OK" if ok else "MISMATCH", out[0].shape, "world", world, flush=True)
# the shared-segment variant (sliced upload + all-gather, every rank writes its block to one host array)
sr = parallel.ShardedReference(x.shape[0], x.shape[1], 300, len(ids), dev)
for _ in range(2):
    res = sr.run(sr.slice_of(x), per, cum, ids)
if rank == 0:
    ok2 = np.array_equal(res[0], idx) and np.array_equal(res[1], dst) and np.allclose(res[2], nr, rtol=1e-13, equal_nan=True)
    print("MGPU_SHM_CHECK", "OK" if ok2 else "MISMATCH", flush=True)
sr.close()
dist.barrier()
dist.destroy_process_group()
