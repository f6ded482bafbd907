"""
Multi-GPU execution of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch)
for the exchanges.  The reference is single-process (no tf.distribute anywhere); the
sharding below follows SURVEY.md 8(e):

* training -- data parallel.  Every rank runs train_step on its own batch (in-batch negatives stay local,
  exactly what one reference process sees for that batch).  Dense gradients: all-gathered and summed over ranks in
  rank order (SUM loss => no 1/G scaling; a step equals the reference's gradient of G independent batches evaluated
  at the same weights and summed, then one optimizer apply).  Embedding tables, two layouts:
    - row-sharded (``shard_tables=True``, the default for world > 1): row i lives on rank i % G.  The forward gather
      reads rows where they live (peer-mapped HBM over NVLink, inside the gather / first-Dense kernels); ranks
      all-gather the batch ids (4 B per id); the owner of a row pulls its gradient rows straight out of the producing
      ranks' dX buffers (peer reads inside the segmented-reduce kernel) and applies the de-duplicated update once.
      Per-rank traffic and optimizer work stay constant as G grows.  The two small NCCL all-gathers (ids, dense
      gradients) also order the ranks: nobody reads a peer's dX before it is complete, nobody gathers a row before its
      owner's previous update has landed.
    - replicated: ranks all-gather (ids, gradient rows) and every replica runs the same update on the whole table.
  Both apply each row's gradients in (feature, rank, position) order, so they are bit-identical to each other and
  to the single-process update of the concatenated batches.
* index -- the corpus is sharded row-wise; each rank returns its shard's exact top-K with GLOBAL row indices,
  ranks all-gather the (score, index) lists and merge them by (score desc, index asc) (tt_topk_merge), so the
  answer does not depend on the number of shards.

The collective helpers are device-agnostic torch.distributed calls, so the world_size-2 gloo tests on CPU
drive exactly this plumbing (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import os
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


class _DevPtr:
    """A raw device address standing in for a tensor in the sparse-source lists (peer-mapped memory has no torch tensor)."""

    def __init__(self, ptr: int, n: int):
        self._ptr, self._n = int(ptr), int(n)

    def data_ptr(self) -> int:
        return self._ptr

    def numel(self) -> int:
        return self._n


class DataParallel:
    """Attach to a compiled-or-not TwoTowerModel: ``DataParallel(model)``; then call model.train_step as usual."""

    def __init__(self, model, group=None, shard_tables: Optional[bool] = None, peer_sync: Optional[bool] = None,
                 global_negatives: bool = False):
        import torch.distributed as dist

        from pkg import _native as N

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        torch = N.require_cuda()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.model = model
        self.shard_tables = (self.world > 1) if shard_tables is None else (bool(shard_tables) and self.world > 1)
        # peer_sync (row-sharded tables only): the ranks are ordered by tt_peer_barrier kernels and read each other's ids / dense
        # gradients in place, so a step contains no NCCL call and is captured as ONE CUDA graph.  False: the two small NCCL
        # all-gathers (ids, dense gradients) order the ranks instead.
        self.peer_sync = self.shard_tables and (os.environ.get("TT_DP_EXCHANGE", "peer") != "nccl" if peer_sync is None else bool(peer_sync))
        # global_negatives (BASELINE configs[4], SURVEY.md 8e "training with global negatives"): every rank scores its B query rows
        # against the candidates of ALL ranks (G.B columns, diagonal offset rank.B) -- the step of ONE reference process on the
        # concatenated batch.  Candidate embeddings / logQ terms are all-gathered and the partial dC reduce-scattered by peer reads
        # between device barriers, so it needs the peer_sync layout.
        self.global_negatives = bool(global_negatives) and self.world > 1
        if self.global_negatives and not self.peer_sync:
            raise ValueError("global_negatives needs row-sharded tables with peer_sync (device barriers + peer reads)")
        model.dist = self
        model._steps.clear()
        # replicas must start identical: the dense parameters are broadcast from rank 0.  Embedding tables nobody has used yet are
        # NOT materialised for that: rank 0 broadcasts their 8-byte initialiser seeds and every rank fills the rows it will own
        # (counter-based initialiser, EmbeddingTable) -- no whole-table allocation or broadcast.  Tables that already hold values
        # (restored from a checkpoint, trained before attaching) are broadcast as before.
        dist.broadcast(model._store.params, src=0, group=group)
        tables = []
        for _, _, t in model._tables():
            if id(t) not in [id(x) for x in tables]:
                tables.append(t)
        meta = torch.tensor([[t.seed, int(t.materialised)] for t in tables], dtype=torch.int64, device="cuda").reshape(-1, 2)
        mine = meta.clone()
        dist.broadcast(meta, src=0, group=group)
        used = torch.maximum(meta[:, 1], mine[:, 1])
        dist.all_reduce(used, op=dist.ReduceOp.MAX, group=group)          # materialised on ANY rank -> take rank 0's values
        for t, (seed, _), u in zip(tables, meta.tolist(), used.tolist()):
            if u or not self.shard_tables:
                dist.broadcast(t.weight, src=0, group=group)
            else:
                t.seed = int(seed)
        if self.shard_tables:
            seen = set()
            for _, _, t in model._tables():
                if id(t) not in seen:
                    seen.add(id(t))
                    t.shard_rows(group)
            for tower in (model.query_tower, model.candidate_tower):
                tower._ws.clear()               # cached descriptors point at the unsharded tables
            if model.optimizer is not None:
                model._build_optimizer_state()  # slots take the shard's shape (fresh state: attach before training)
            self.barrier()

    def barrier(self):
        """All ranks' queued work is complete (call before reading tables outside train_step, e.g. to build an index)."""
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    # ---- buffers ---------------------------------------------------------------------------------------
    # Two collectives per step and nothing else outside the two captured compute phases:
    #   (1) ONE all-gather of the id columns.  The towers' id buffers ARE rows of ``ids_local`` (staging writes them in place).
    #   (2) ONE all-gather of every rank's message [dX of the query tower | dX of the candidate tower | flat dense gradients].
    #       The towers' dX buffers ARE slices of ``msg_local`` (the tower backward writes them in place); the dense gradients are
    #       copied in at the end of phase A and summed over ranks in rank order at the start of phase B (a fixed order: replicas
    #       stay bit-identical) -- both inside the captured graphs.
    # The sparse optimizer reads the gathered buffers where NCCL put them: one source per (feature, rank), in (feature, rank,
    # position) order -- the order the single-process update of the concatenated batches would use.
    def _ensure(self, sw):
        if getattr(sw, "dp", None) is not None:
            return sw.dp
        from pkg import _native as N
        from pkg.modelling import _device as D

        torch = N.require_cuda()
        g, b = self.world, sw.batch
        model = self.model
        towers = [(model.query_tower, sw.q), (model.candidate_tower, sw.c)]
        id_slots = []                                   # (tower index, block index) of every id column
        for ti, (tower, tws) in enumerate(towers):
            for bi, (_, t, _, _) in enumerate(tower.input_layer.blocks):
                if t is not None:
                    id_slots.append((ti, bi))
        n_id = len(id_slots)
        lds = [tower.input_layer.ld for tower, _ in towers]
        n_dense = int(model._store.used)
        offs = [0, b * lds[0]]                          # first float of each tower's dX block inside a message
        dense_off = b * (lds[0] + lds[1])
        dp = {
            "towers": towers, "id_slots": id_slots, "lds": lds, "offs": offs, "dense_off": dense_off, "n_dense": n_dense,
            "ids_local": torch.zeros((max(n_id, 1), b), dtype=torch.int32, device="cuda"),
            "ids_all": torch.zeros((g, max(n_id, 1), b), dtype=torch.int32, device="cuda"),      # [rank, column, position]
        }
        if self.shard_tables:
            # the dX blocks live in a peer-shareable buffer (read in place by the rows' owners); only the dense gradients travel
            from pkg.modelling._peer import PeerBuffer

            dp["peer_msg"] = PeerBuffer((max(dense_off, 64),), "float32", self.group)     # collective: same batch on every rank
            dp["msg_local"] = dp["peer_msg"].local
            dp["grad_base"] = list(dp["peer_msg"].ptrs)
            if self.peer_sync:
                dp["peer_ids"] = PeerBuffer((max(n_id, 1), b), "int32", self.group)
                dp["ids_local"] = dp["peer_ids"].local
                dp["peer_dense"] = PeerBuffer((max(n_dense, 64),), "float32", self.group)
                dp["dense_local"] = dp["peer_dense"].local
                dp["peer_flags"] = PeerBuffer((N.TT_PEER_SLOTS * (1 + g),), "int32", self.group)
                if self.global_negatives:
                    e = model.joint_embedding_size
                    if b % 4:
                        raise ValueError("global_negatives: the per-rank batch must be a multiple of 4")
                    lib = N.load()
                    use_tc = model.impl != N.TT_IMPL_SIMT and model._tc_ok()
                    dp["peer_c"] = PeerBuffer((b, e), "float32", self.group)          # this rank's candidate embeddings, as the softmax consumes them
                    if use_tc:
                        sw.c.out_tf32 = dp["peer_c"].local
                    else:
                        sw.c.acts[-1] = dp["peer_c"].local
                    dp["c_all"] = torch.zeros((g * b, e), dtype=torch.float32, device="cuda")
                    if sw.col_bias is not None:
                        dp["peer_bias"] = PeerBuffer((b,), "float32", self.group)
                        sw.col_bias = dp["peer_bias"].local
                        dp["bias_all"] = torch.zeros(g * b, dtype=torch.float32, device="cuda")
                    dp["peer_dc"] = PeerBuffer((g * b, e), "float32", self.group)     # d(loss of MY rows)/d(all candidates)
                    # reduce-scatter sources: every rank's partial, offset to the slice of my own candidates
                    dp["dc_srcs"] = torch.tensor([p + 4 * self.rank * b * e for p in dp["peer_dc"].ptrs], dtype=torch.int64, device="cuda")
                    sw.sm_ws = torch.empty(int(lib.tt_softmax_workspace_bytes(b, g * b, e)), dtype=torch.uint8, device="cuda")
                self.barrier()                          # every rank's (zeroed) flag block exists before the first device barrier
            else:
                dp["dense_local"] = torch.zeros(max(n_dense, 1), dtype=torch.float32, device="cuda")
                dp["dense_all"] = torch.zeros((g, max(n_dense, 1)), dtype=torch.float32, device="cuda")
        else:
            msg = (dense_off + n_dense + 63) // 64 * 64     # floats per rank (every rank's block stays 256-byte aligned)
            if g * msg * 4 >= 2 ** 40:
                raise ValueError("data-parallel message too large")
            dp["msg_local"] = torch.zeros(msg, dtype=torch.float32, device="cuda")
            dp["msg_all"] = torch.zeros((g, msg), dtype=torch.float32, device="cuda")
            dp["grad_base"] = [dp["msg_all"][r].data_ptr() for r in range(g)]
            dp["dense_local"] = dp["msg_local"][dense_off:dense_off + n_dense]
            dp["dense_all"] = dp["msg_all"][:, dense_off:dense_off + n_dense]
        # alias the towers' staging / gradient buffers into the exchange buffers (no packing copies per step)
        for k, (ti, bi) in enumerate(id_slots):
            towers[ti][1].bufs[bi] = dp["ids_local"][k]
        for ti, (tower, tws) in enumerate(towers):
            tws.feats = tower.input_layer.descriptors(tws.bufs)
            tws.dx = dp["msg_local"][offs[ti]:offs[ti] + b * lds[ti]].view(b, lds[ti])
        if model.logq_correction is not None:           # the ln p(candidate) gather reads the candidate-id column
            for (f, t, _, _), buf in zip(model.candidate_tower.input_layer.blocks, sw.c.bufs):
                if f.name == model.candidate_id_col and t is not None:
                    sw.cid_buf = buf
            sw.bias_feat = D.feature_array([{"table": model._logq_rows.data_ptr(), "src": sw.cid_buf.data_ptr(),
                                             "rows": model._logq_rows.shape[0], "e": 1, "col": 0}])
        sw.dp = dp
        return dp

    def build_sparse_sources(self, model, sw):
        """Same structure as the single-GPU source list, but every feature contributes one source per rank, read straight from
        the all-gathered buffers."""
        from pkg import _native as N

        dp = self._ensure(sw)
        g = self.world
        srcs = []
        slot = 0
        for ti, (tower, tws) in enumerate(dp["towers"]):
            il = tower.input_layer
            per_table = {}
            for bi, (f, t, col, w) in enumerate(il.blocks):
                if t is None:
                    continue
                lst = per_table.setdefault(f.name, (t, []))[1]
                for r in range(g):
                    # peer_sync: rank r's id column is read in place from its (peer-mapped) staging buffer
                    ids = _DevPtr(dp["peer_ids"].ptrs[r] + 4 * slot * sw.batch, sw.batch) if self.peer_sync else dp["ids_all"][r, slot]
                    lst.append((ids, dp["grad_base"][r] + 4 * (dp["offs"][ti] + col), dp["lds"][ti]))
                slot += 1
            for name, (t, lst) in per_table.items():
                if len(lst) > N.TT_MAX_SRC:
                    raise ValueError(f"table {name}: {len(lst) // g} features x {g} ranks exceed {N.TT_MAX_SRC} gradient sources")
            srcs.extend(per_table.values())
        return srcs

    # ---- exchanges (called from TwoTowerModel.train_step) ------------------------------------------------
    def _device_barrier(self, dp, slot):
        from pkg import _native as N

        N.check(N.load().tt_peer_barrier(dp["peer_flags"].ptr_table.data_ptr(), self.rank, self.world, slot, N.stream_ptr()), "tt_peer_barrier")

    def gather_ids(self, model, sw):
        """Start of a step.  Every rank's ids of this step are staged and its previous table update has landed: either the id
        all-gather says so (NCCL), or a device barrier does and the ids are read where they are (peer_sync)."""
        dp = self._ensure(sw)
        if self.peer_sync:
            self._device_barrier(dp, 0)
        else:
            allgather_into(dp["ids_all"].view(self.world * dp["ids_local"].shape[0], -1), dp["ids_local"], self.group)

    def gather_candidates(self, model, sw):
        """global_negatives: (C_all, ln p_all or None, diagonal offset, number of columns) once every rank's candidate tower is done."""
        from pkg import _native as N

        dp = self._ensure(sw)
        lib = N.load()
        b, e, st = sw.batch, model.joint_embedding_size, N.stream_ptr()
        self._device_barrier(dp, 2)
        N.check(lib.tt_peer_gather_f32(dp["peer_c"].ptr_table.data_ptr(), self.world, b * e, dp["c_all"].data_ptr(), st), "tt_peer_gather_f32(C)")
        bias_all = None
        if "peer_bias" in dp:
            N.check(lib.tt_peer_gather_f32(dp["peer_bias"].ptr_table.data_ptr(), self.world, b, dp["bias_all"].data_ptr(), st), "tt_peer_gather_f32(ln p)")
            bias_all = dp["bias_all"]
        return dp["c_all"], bias_all, self.rank * b, self.world * b

    def reduce_dc(self, model, sw):
        """global_negatives: dC of my own candidates = sum over ranks (rank order) of their partial gradients for my slice."""
        from pkg import _native as N

        dp = self._ensure(sw)
        self._device_barrier(dp, 3)
        N.check(N.load().tt_peer_sum_f32(dp["dc_srcs"].data_ptr(), self.world, sw.batch * model.joint_embedding_size, sw.dc.data_ptr(),
                                         N.stream_ptr()), "tt_peer_sum_f32(dC)")

    def pack_dense(self, model, sw):
        """End of phase A (captured): the flat dense gradients join the message."""
        dp = self._ensure(sw)
        if dp["n_dense"]:
            dp["dense_local"][:dp["n_dense"]].copy_(model._store.grads[:dp["n_dense"]])

    def gather_rows(self, model, sw):
        """The one exchange between the compute phases: [dX blocks | dense gradients] (replicated tables) or the dense gradients
        alone (row-sharded tables; the all-gather then also tells every rank that all dX blocks are complete)."""
        dp = self._ensure(sw)
        if self.peer_sync:
            self._device_barrier(dp, 1)                 # every rank's dX blocks and dense gradients are complete
        elif self.shard_tables:
            allgather_into(dp["dense_all"], dp["dense_local"], self.group)
        else:
            allgather_into(dp["msg_all"], dp["msg_local"], self.group)

    def sum_dense(self, model, sw):
        """Start of phase B (captured): dense gradients summed over ranks in rank order."""
        import torch

        dp = self._ensure(sw)
        if dp["n_dense"] and self.peer_sync:
            from pkg import _native as N

            N.check(N.load().tt_peer_sum_f32(dp["peer_dense"].ptr_table.data_ptr(), self.world, dp["n_dense"], model._store.grads.data_ptr(),
                                             N.stream_ptr()), "tt_peer_sum_f32")
        elif dp["n_dense"]:
            torch.sum(dp["dense_all"][:, :dp["n_dense"]], dim=0, out=model._store.grads[:dp["n_dense"]])


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
