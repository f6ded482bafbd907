#!/usr/bin/env python
"""torchrun diagnostic: which parameter tensors differ between the N-rank cross-GPU-negatives step and the single-GPU step on the
concatenated batch (bench.py's `parity.global_negatives`), per tensor."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from pkg import _native as N  # noqa: E402
from pkg.modelling.distributed import DataParallel  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
impl = int(os.environ.get("TT_DIAG_IMPL", "0"))
whole = {k: torch.from_numpy(np.concatenate([bench._par_batch(r)[k] for r in range(world)], axis=0)).cuda() for k in bench._par_batch(0)}
ref = bench._small_model(78, 0.05, 0.1); ref.impl = impl
w0 = {k: v.copy() for k, v in ref.state_arrays().items()}
loss_1 = float(ref.train_step(whole)["loss"])
want = ref.state_arrays()
gn = bench._small_model(78, 0.05, 0.1); gn.impl = impl
DataParallel(gn, shard_tables=True, global_negatives=True)
mine = {k: torch.from_numpy(v).cuda() for k, v in bench._par_batch(rank).items()}
loss_g = float(gn.train_step(mine)["loss"])
gn.dist.barrier()
got = gn.state_arrays()
t = torch.tensor([loss_g], device="cuda", dtype=torch.float64); dist.all_reduce(t)
if rank == 0:
    print(f"world {world} impl {impl}: loss single {loss_1:.4f}  sum over ranks {float(t):.4f}")
    for k in want:
        upd = np.abs(want[k] - w0[k]).max()
        err = np.abs(got[k] - want[k])
        i = np.unravel_index(err.argmax(), err.shape)
        print(f"  {k:45s} max|update| {upd:.3e}  max err {err.max():.3e} at {i}: got {got[k][i]:.6f} want {want[k][i]:.6f} was {w0[k][i]:.6f}   rows with err>1e-4: {int((err.max(axis=-1) > 1e-4).sum()) if err.ndim == 2 else '-'}")
dist.destroy_process_group()
