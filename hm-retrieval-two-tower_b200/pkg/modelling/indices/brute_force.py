"""
BruteForceIndex: keep (ids, candidate embeddings), answer the top-k ids per query
(reference pkg/modelling/indices/brute_force.py:7-114).

call = query_model(x) -> scores = q.C^T -> top_k (sorted descending, LOWER index first on ties) -> ids.
Scoring and selection are one fused CUDA path (tt_index_topk): the (B x N) score matrix is never written
to HBM.  Indices and scores are bit-identical to the oracle's canonical fp32 evaluation.  The final
index -> identifier gather stays on the host (identifiers are strings in the reference, :83).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from pkg import _native as N
from pkg.modelling.models.abstract_keras_model import AbstractKerasModel, TensorSpec


def _to_device_f32(a):
    torch = N.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).cuda()


class BruteForceIndex(AbstractKerasModel):
    def __init__(self, k: int, query_model, id_candidate_pairs: Iterable, shard: Optional[Tuple[int, int]] = None,
                 _local_rows: Optional[Tuple[int, int]] = None):
        """``id_candidate_pairs`` yields (ids (n,), embeddings (n, E)) batches.  ``shard=(rank, world)`` keeps only
        this rank's contiguous slice of rows (see pkg.modelling.distributed.make_sharded_index)."""
        super().__init__()
        self.k = int(k)
        self.query_model = query_model
        self.impl = N.TT_IMPL_AUTO
        self._shard = shard
        self._index(id_candidate_pairs, _local_rows)
        self._ws = None
        self._pos_of = None
        self.initialise_model()

    @classmethod
    def from_local_rows(cls, k: int, query_model, candidates_local, row_base: int, n_total: int, identifiers=None, group=None,
                        world: int = 1) -> "BruteForceIndex":
        """Row-sharded corpus whose shards are built where they live: this rank hands over ONLY its own rows
        ``[row_base, row_base + n_local)`` of an ``n_total``-row corpus (no rank ever holds the whole corpus).
        ``identifiers``: all ``n_total`` identifiers (position == global row), or None when a row's identifier is its
        global row number.  ``world`` > 1: results are merged over the ranks of ``group`` like make_sharded_index's."""
        n_local = int(candidates_local.shape[0])
        if row_base < 0 or row_base + n_local > n_total:
            raise ValueError(f"rows [{row_base}, {row_base + n_local}) lie outside a corpus of {n_total} rows")
        if identifiers is not None and len(identifiers) != n_total:
            raise ValueError(f"{len(identifiers)} identifiers for a corpus of {n_total} rows")
        index = cls(k, query_model, [(identifiers, candidates_local)], _local_rows=(int(row_base), int(n_total)))
        index._group, index._world = group, int(world)
        return index

    def _index(self, id_candidate_pairs: Iterable, local_rows: Optional[Tuple[int, int]] = None) -> None:
        torch = N.require_cuda()
        if local_rows is not None:
            (identifiers, candidates), = list(id_candidate_pairs)
            candidates = _to_device_f32(candidates)
            self.idx_base, n_total = local_rows
            if identifiers is not None:
                identifiers = np.asarray(identifiers).reshape(-1)
        else:
            identifiers, candidates = self.get_id_embeddings_from_dataset(id_candidate_pairs)
            n_total = candidates.shape[0]
            self.idx_base = 0
            if self._shard is not None:
                rank, world = self._shard
                per = (n_total + world - 1) // world
                lo, hi = min(n_total, rank * per), min(n_total, (rank + 1) * per)
                self.idx_base = lo
                candidates = candidates[lo:hi]
        self.n_total = n_total
        self._identifiers = identifiers                       # all ids (host), position == global row index; None: id == row number
        # numeric identifiers are also kept on the device: `call` then maps row indices to identifiers there and
        # copies only the (B, k) result out (through a pinned buffer) instead of gathering on the host
        self._identifiers_dev = (torch.from_numpy(np.ascontiguousarray(identifiers)).cuda()
                                 if identifiers is not None and np.issubdtype(np.asarray(identifiers).dtype, np.integer) else None)
        self._candidates = candidates.contiguous()            # (N_local, E) fp32, non-trainable
        # operand preparation for the tensor-core filter, done once at build time: permuted TF32-rounded copy of
        # the corpus and the row norms (error bound of the filter); the exact fp32 rows stay authoritative
        lib = N.load()
        n_loc, e = self._candidates.shape
        self._candidates_tf32, self._max_norm = None, None
        if n_loc > 0 and lib.tt_tc_available(1, e):
            # opaque operand tiles (fp16 from E = 64: half the size of the corpus); named for the round-1 TF32 layout
            self._candidates_tf32 = torch.empty(int(lib.tt_index_prepared_bytes(n_loc, e)), dtype=torch.uint8, device="cuda")
            rows_pad = ((n_loc + 255) // 256 + 1) * 256          # TT_INDEX_ROWS_PAD
            n_pad = 2 * rows_pad + rows_pad // 32 + 32           # TT_INDEX_NORM_PAD: row norms + per-chunk maxima + row map + scale slot
            self._max_norm = torch.zeros(n_pad, dtype=torch.float32, device="cuda")   # per-row norms, zero padded
            N.check(lib.tt_index_prepare(self._candidates.data_ptr(), e, n_loc, e, self._candidates_tf32.data_ptr(),
                                         self._max_norm.data_ptr(), N.stream_ptr()), "tt_index_prepare")
        if self.k > n_total:
            raise ValueError(f"k={self.k} exceeds the number of candidates ({n_total})")

    @staticmethod
    def get_id_embeddings_from_dataset(candidates: Iterable):
        """Concatenate an iterable of (ids, embeddings) -> (ids (N,), embeddings (N, E) on the device)."""
        torch = N.require_cuda()
        ids, embs = [], []
        for identifiers, embeddings in candidates:
            if isinstance(identifiers, torch.Tensor):
                identifiers = identifiers.detach().cpu().numpy()
            ids.append(np.asarray(identifiers).reshape(-1))
            e = _to_device_f32(embeddings)
            embs.append(e.reshape(-1, e.shape[-1]))
        if not embs:
            raise ValueError("id_candidate_pairs is empty")
        return np.concatenate(ids, axis=0), torch.cat(embs, dim=0)

    # ---- query paths -----------------------------------------------------------------------------------
    def _embed_queries(self, queries):
        torch = N.require_cuda()
        out = self.query_model(queries)
        return _to_device_f32(out).contiguous()

    def search(self, query_embeddings, k: Optional[int] = None):
        """(B, E) device embeddings -> (scores (B,k) fp32, GLOBAL row indices (B,k) int32), both on the device."""
        torch = N.require_cuda()
        lib = N.load()
        k = int(k or self.k)
        q = query_embeddings
        nq, e = q.shape
        n = self._candidates.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        idx = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        if n == 0 or nq == 0:
            scores.fill_(float("-inf"))
            idx.fill_(-1)
            if getattr(self, "_world", 1) > 1 and nq > 0:
                from pkg.modelling.distributed import merge_shard_results

                scores, idx = merge_shard_results(scores, idx, k, getattr(self, "_group", None))
            return scores, idx
        have32 = self._candidates_tf32 is not None
        need = int(lib.tt_index_workspace_bytes(nq, n, e, k, self.impl, 1 if have32 else 0))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        N.check(lib.tt_index_topk(q.data_ptr(), q.stride(0), self._candidates.data_ptr(), self._candidates.stride(0),
                                  self._candidates_tf32.data_ptr() if have32 else None, self._max_norm.data_ptr() if have32 else None,
                                  nq, n, e, k, self.idx_base, scores.data_ptr(), idx.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                  self.impl, N.stream_ptr()), "tt_index_topk")
        if getattr(self, "_world", 1) > 1:   # row-sharded corpus: merge the per-shard lists (identical on every rank)
            from pkg.modelling.distributed import merge_shard_results

            scores, idx = merge_shard_results(scores, idx, k, getattr(self, "_group", None))
        return scores, idx

    def query_indices(self, queries, k: Optional[int] = None):
        return self.search(self._embed_queries(queries), k)

    def call(self, queries, training: bool = False, out=None, wait: bool = True):
        """{query feature: (B,1)} -> (B, k) array of candidate identifiers.

        `out` (optional): a pinned host tensor of shape (B, k) and the identifiers' dtype that receives the result; the returned numpy
        array then views it (no further host copy -- the caller owns the buffer and decides when it is reused).  Without `out` the
        result is an array of its own.  `wait=False` (needs `out`): do not wait for the device -- returns (view of `out`, CUDA event);
        the view holds the result once ``event.synchronize()`` returns, so a serving loop can submit batch k+1 before reading batch k."""
        torch = N.require_cuda()
        lib = N.load()
        _, idx = self.query_indices(queries)
        if self._identifiers_dev is not None or self._identifiers is None:
            if self._identifiers is None:
                ids_dev = idx
            elif self._identifiers_dev.dtype == torch.int32:
                ids_dev = torch.empty_like(idx)
                N.check(lib.tt_take_i32(self._identifiers_dev.data_ptr(), idx.data_ptr(), idx.numel(), ids_dev.data_ptr(), N.stream_ptr()), "tt_take_i32")
            else:
                ids_dev = self._identifiers_dev[idx.long().clamp_(min=0)]
            shape = tuple(ids_dev.shape)
            if out is not None:
                if tuple(out.shape) != shape or out.dtype != ids_dev.dtype or not out.is_pinned():
                    raise ValueError(f"out must be a pinned host tensor of shape {shape} and dtype {ids_dev.dtype}")
                out.copy_(ids_dev, non_blocking=True)
                if not wait:
                    done = torch.cuda.Event()
                    done.record()
                    return out.numpy(), done
                torch.cuda.current_stream().synchronize()
                return out.numpy()
            if not wait:
                raise ValueError("wait=False needs a caller-owned pinned `out` buffer")
            stage = self.__dict__.get("_pin_stage")
            if stage is None or tuple(stage.shape) != shape or stage.dtype != ids_dev.dtype:
                stage = self.__dict__["_pin_stage"] = torch.empty(shape, dtype=ids_dev.dtype).pin_memory()
            stage.copy_(ids_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return stage.numpy().copy()    # the caller's own array: the staging buffer is reused by the next call
        return self._identifiers[idx.cpu().numpy()]

    def canonical_rows(self):
        """None when every identifier occurs once; else a device int32 array row -> FIRST row holding the same identifier (the row
        positions_of reports), so that Recall@k on row indices counts every row of a repeated identifier like the reference's
        comparison of identifiers does (index_recall.py:54-58)."""
        if "_canon" not in self.__dict__:
            canon = None
            if self._identifiers is not None:
                ids = np.asarray(self._identifiers)
                keys = ids if ids.dtype.kind in "iu" else np.array([v.decode() if isinstance(v, bytes) else str(v) for v in ids.reshape(-1)])
                _, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
                if first.shape[0] < keys.shape[0]:
                    torch = N.require_cuda()
                    canon = torch.from_numpy(first[inverse].astype(np.int32)).cuda()
            self.__dict__["_canon"] = canon
        return self.__dict__["_canon"]

    def positions_of(self, ids) -> np.ndarray:
        """Global row index of each identifier (-1 when absent); used by IndexRecall's device path."""
        flat = np.asarray(ids).reshape(-1)
        if self._identifiers is None:     # identifier == global row number
            pos = np.array([int(v) for v in flat], dtype=np.int64)
            return np.where((pos >= 0) & (pos < self.n_total), pos, -1).astype(np.int32)
        if self._pos_of is None:
            self._pos_of = {}
            for i, v in enumerate(self._identifiers):
                self._pos_of.setdefault(v.decode() if isinstance(v, bytes) else str(v), i)
        return np.fromiter((self._pos_of.get(v.decode() if isinstance(v, bytes) else str(v), -1) for v in flat), dtype=np.int32,
                           count=flat.shape[0])

    def get_input_signature(self) -> Dict[str, TensorSpec]:
        return self.query_model.get_input_signature()

    # ---- bulk build / reload -----------------------------------------------------------
    @classmethod
    def from_candidate_tower(cls, k: int, query_model, candidate_tower, candidate_batches: Iterable, id_col: str,
                             shard: Optional[Tuple[int, int]] = None) -> "BruteForceIndex":
        """The reference's ``candidate_ds.map(lambda x: (x[id_col], model.candidate_tower(x)))`` (runner.py:88-93) without a host
        round trip: every batch is embedded on the device straight into one pre-sized (N, E) corpus buffer (no per-batch
        tensors, no concatenate), which the index then adopts and prepares."""
        torch = N.require_cuda()
        batches = list(candidate_batches)
        sizes = [int(np.asarray(b[id_col]).reshape(-1).shape[0]) for b in batches]
        e = candidate_tower.joint_embedding_size
        corpus = torch.empty((sum(sizes), e), dtype=torch.float32, device="cuda")
        ids, lo = [], 0
        for b, n in zip(batches, sizes):
            out, _ = candidate_tower.embed_tf32({f.name: b[f.name] for f in candidate_tower.features})
            corpus[lo:lo + n].copy_(out)
            ids.append(np.asarray(b[id_col]).reshape(-1))
            lo += n
        return cls(k, query_model, [(np.concatenate(ids, axis=0), corpus)], shard=shard)

    def save(self, model_path: str) -> None:
        """identifiers + exact corpus rows as variables.npz under ``model_path`` (the reference exports a SavedModel here)."""
        super().save(model_path)

    @classmethod
    def load(cls, model_path: str, k: int, query_model, shard: Optional[Tuple[int, int]] = None) -> "BruteForceIndex":
        """Rebuild an index from what ``save`` wrote; the prepared (tensor-core) copy is regenerated on the device."""
        import os

        f = model_path if model_path.endswith(".npz") else os.path.join(model_path, "variables.npz")
        with np.load(f) as z:
            ids, cand = z["identifiers"], z["candidates"]
        return cls(k, query_model, [(ids, cand)], shard=shard)

    def state_arrays(self) -> Dict[str, np.ndarray]:
        ids = np.asarray(self._identifiers if self._identifiers is not None else np.arange(self.n_total, dtype=np.int32))
        return {"identifiers": ids if np.issubdtype(ids.dtype, np.integer) else ids.astype(str),
                "candidates": self._candidates.detach().cpu().numpy()}
