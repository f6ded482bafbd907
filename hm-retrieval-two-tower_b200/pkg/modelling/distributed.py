"""
Multi-GPU execution of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch)
for the exchanges.  The reference is single-process (no tf.distribute anywhere); the
sharding below follows SURVEY.md 8(e):

* training -- data parallel.  Every rank runs train_step on its own batch (in-batch negatives stay local,
  exactly what one reference process sees for that batch).  Dense gradients: ONE all-reduce(SUM) of the flat
  gradient buffer (SUM loss => no 1/G scaling; a step equals the reference's gradient of G independent
  batches evaluated at the same weights and summed, then one optimizer apply).  Embedding tables are
  replicated: ranks all-gather (ids, gradient rows) and every replica runs the same deterministic
  de-duplicated row update, ordered by (rank, position), so replicas stay bit-identical.
* index -- the corpus is sharded row-wise; each rank returns its shard's exact top-K with GLOBAL row indices,
  ranks all-gather the (score, index) lists and merge them by (score desc, index asc) (tt_topk_merge), so the
  answer does not depend on the number of shards.

The collective helpers are device-agnostic torch.distributed calls, so the world_size-2 gloo tests on CPU
drive exactly this plumbing (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of shard ``rank`` (the last shards may be short or empty)."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def allreduce_sum_(t, group=None):
    """In-place SUM all-reduce of a flat tensor (dense Dense-layer gradients)."""
    import torch.distributed as dist

    if dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_into(out, local, group=None):
    """out (world*n, ...) <- concatenation over ranks of local (n, ...), rank-major: the (rank, position)
    order the deterministic sparse update relies on."""
    import torch.distributed as dist

    if dist.get_world_size(group) == 1:
        out.copy_(local)
    elif hasattr(dist, "all_gather_into_tensor") and local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:   # gloo: list form
        world = dist.get_world_size(group)
        chunks = list(out.view(world, *local.shape).unbind(0))
        dist.all_gather(chunks, local.contiguous(), group=group)
    return out


class DataParallel:
    """Attach to a compiled-or-not TwoTowerModel: ``DataParallel(model)``; then call model.train_step as usual."""

    def __init__(self, model, group=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.model = model
        model.dist = self
        model._steps.clear()
        # replicas must start identical: parameters are broadcast from rank 0
        dist.broadcast(model._store.params, src=0, group=group)
        for _, _, t in model._tables():
            dist.broadcast(t.weight, src=0, group=group)

    # ---- buffers ---------------------------------------------------------------------------------------
    def _ensure(self, sw):
        if getattr(sw, "dp", None) is not None:
            return sw.dp
        from pkg import _native as N

        torch = N.require_cuda()
        g, b = self.world, sw.batch
        dp = {"towers": []}
        for tower, tws in ((self.model.query_tower, sw.q), (self.model.candidate_tower, sw.c)):
            il = tower.input_layer
            ids_all = [None if t is None else torch.zeros(g * b, dtype=torch.int32, device="cuda") for (_, t, _, _) in il.blocks]
            dx_all = torch.zeros((g * b, il.ld), dtype=torch.float32, device="cuda")
            dp["towers"].append((tower, tws, ids_all, dx_all))
        sw.dp = dp
        return dp

    def build_sparse_sources(self, model, sw):
        """Same structure as the single-GPU source list, but over the all-gathered buffers."""
        dp = self._ensure(sw)
        srcs = []
        for tower, tws, ids_all, dx_all in dp["towers"]:
            il = tower.input_layer
            per_table = {}
            for (f, t, col, w), ga in zip(il.blocks, ids_all):
                if t is None:
                    continue
                per_table.setdefault(f.name, (t, []))[1].append((ga, dx_all.data_ptr() + 4 * col, il.ld))
            srcs.extend(per_table.values())
        return srcs

    # ---- exchanges (called from TwoTowerModel._launch_step) ---------------------------------------------
    def gather_ids(self, model, sw):
        dp = self._ensure(sw)
        for tower, tws, ids_all, _ in dp["towers"]:
            for buf, ga in zip(tws.bufs, ids_all):
                if ga is not None:
                    allgather_into(ga, buf, self.group)

    def reduce_dense_and_gather_rows(self, model, sw):
        dp = self._ensure(sw)
        used = model._store.used
        if used:
            allreduce_sum_(model._store.grads[:used], self.group)
        for tower, tws, _, dx_all in dp["towers"]:
            allgather_into(dx_all, tws.dx, self.group)


def make_sharded_index(k: int, query_model, id_candidate_pairs: Iterable, group=None):
    """BruteForceIndex whose corpus rows are split over the ranks of ``group``; queries are replicated."""
    import torch.distributed as dist

    from pkg.modelling.indices.brute_force import BruteForceIndex

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    index = BruteForceIndex(k, query_model, id_candidate_pairs, shard=(rank, world))
    index._group = group
    index._world = world
    return index


def merge_shard_results(scores, idx, k: int, group=None):
    """All-gather per-shard (nq, k) results and merge them on the device (identical on every rank)."""
    import torch.distributed as dist

    from pkg import _native as N

    torch = N.require_cuda()
    lib = N.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return scores, idx
    nq = scores.shape[0]
    all_s = torch.empty((world * nq, k), dtype=torch.float32, device="cuda")
    all_i = torch.empty((world * nq, k), dtype=torch.int32, device="cuda")
    allgather_into(all_s, scores.contiguous(), group)
    allgather_into(all_i, idx.contiguous(), group)
    out_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    N.check(lib.tt_topk_merge(all_s.data_ptr(), all_i.data_ptr(), world, nq, k, out_s.data_ptr(), out_i.data_ptr(), N.stream_ptr()),
            "tt_topk_merge")
    return out_s, out_i
