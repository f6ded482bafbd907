"""
IndexRecall: Recall@k of an index over batches of (queries, true candidate id)
(reference pkg/modelling/metrics/index_recall.py:12-84).

hits[k] += #{(b, j < k) : true_b == cand[b, j]},  seen += B,  metric[k] = hits[k] / seen  (int32 / int32 ->
float64).  With a BruteForceIndex the comparison runs on the device on row indices (tt_recall_hits,
integer-exact); any other index (StaticIndex, user objects returning id arrays) is compared on the host
exactly as the reference does.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np

from pkg import _native as N
from pkg.modelling import _device as D

logger = logging.getLogger(__name__)
_SUMMARY_WRITER = [None]


def set_summary_writer(writer) -> None:
    """Optional torch.utils.tensorboard SummaryWriter receiving "Epoch Start Recall@k" scalars."""
    _SUMMARY_WRITER[0] = writer


class IndexRecall:
    def __init__(self, index, ks: List[int]):
        self.index = index
        self.ks = list(ks)
        if len(self.ks) > N.TT_MAX_KS:
            raise ValueError(f"at most {N.TT_MAX_KS} cut-offs")
        # hits / seen / metric are what the reference exposes.  This process's own counts live in _local_*: all_reduce() sums THOSE
        # over the ranks and publishes the totals in hits / seen / metric, so calling it twice (or scoring more batches and
        # calling it again) never counts anything twice.
        self.hits = {k: np.int32(0) for k in self.ks}
        self.seen = np.int32(0)
        self.metric = {k: np.int32(0) for k in self.ks}
        self._local_hits = {k: np.int32(0) for k in self.ks}
        self._local_seen = np.int32(0)
        self._dev_hits = None
        self._dev_ks = None

    def _device_update(self, queries, true_candidate_ids) -> None:
        torch = N.require_cuda()
        lib = N.load()
        _, idx = self.index.query_indices(queries)
        canon = self.index.canonical_rows() if hasattr(self.index, "canonical_rows") else None
        if canon is not None:      # repeated identifiers: the reference compares IDENTIFIERS (tf.equal), so every row of an id counts
            idx = torch.where(idx >= 0, canon[idx.clamp(min=0).long()], idx).contiguous()
        if isinstance(true_candidate_ids, torch.Tensor) and true_candidate_ids.dtype in (torch.int32, torch.int64) \
                and getattr(self.index, "identifiers_are_positions", False):
            truth = true_candidate_ids.reshape(-1).to(device="cuda", dtype=torch.int32)
        else:
            t = true_candidate_ids.detach().cpu().numpy() if isinstance(true_candidate_ids, torch.Tensor) else true_candidate_ids
            truth = torch.from_numpy(self.index.positions_of(t)).cuda()
        if self._dev_hits is None:
            self._dev_hits = torch.zeros(len(self.ks), dtype=torch.int32, device="cuda")
            self._dev_ks = np.asarray(self.ks, dtype=np.int32)
        N.check(lib.tt_recall_hits(idx.data_ptr(), idx.stride(0), truth.data_ptr(), idx.shape[0],
                                   self._dev_ks.ctypes.data, len(self.ks), self._dev_hits.data_ptr(), N.stream_ptr()), "tt_recall_hits")
        host = self._dev_hits.cpu().numpy()
        for k, h in zip(self.ks, host):
            self._local_hits[k] = np.int32(h)

    def __call__(self, queries, true_candidate_ids) -> Dict[int, float]:
        true_candidate_ids = D.unwrap(true_candidate_ids)      # TF / DLPack tensors -> torch or numpy
        n = int(true_candidate_ids.shape[0])
        self._local_seen = np.int32(self._local_seen + n)
        if hasattr(self.index, "query_indices"):
            self._device_update(queries, true_candidate_ids)
        else:
            candidates = np.asarray(self.index(queries))
            truth = np.asarray(true_candidate_ids).reshape(-1, 1)
            if truth.dtype.kind in ("S", "O", "U") or candidates.dtype.kind in ("S", "O", "U"):
                # ids may arrive as str on one side and bytes on the other (TFRecords yield bytes).  The (B, 1) truth column is
                # brought to the dtype of the (B, k) candidate block, which is then compared in one vectorised pass -- not a Python
                # call per element: the reference's main.py evaluates k = 1000 candidates per query
                if candidates.dtype.kind == "O":
                    candidates = D._as_bytes_array(candidates).reshape(candidates.shape)
                if candidates.dtype.kind == "U":
                    flat = truth.reshape(-1)
                    truth = np.array([D._as_str(v) for v in flat]).reshape(-1, 1) if flat.shape[0] else truth.astype("U1")
                else:
                    truth = D._as_bytes_array(truth).reshape(-1, 1)
            for k in self.ks:
                self._local_hits[k] = np.int32(self._local_hits[k] + np.int32(np.sum(truth == candidates[:, :k])))
        self.seen = self._local_seen
        for k in self.ks:
            self.hits[k] = self._local_hits[k]
            self.metric[k] = np.float64(self.hits[k]) / np.float64(self.seen)
        return self.metric

    def all_reduce(self, group=None) -> Dict[int, float]:
        """Multi-GPU evaluation: every rank has scored its own share of the test queries; the int32 hit and seen
        counters are summed over the ranks (exact) and the ratios recomputed, so every rank ends with the global Recall@k in
        hits / seen / metric.  Idempotent: the ranks' own counts are kept apart and are what is summed."""
        import torch
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self.metric
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor([int(self._local_hits[k]) for k in self.ks] + [int(self._local_seen)], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        vals = t.cpu().numpy()
        for k, h in zip(self.ks, vals[:-1]):
            self.hits[k] = np.int32(h)
        self.seen = np.int32(vals[-1])
        for k in self.ks:
            self.metric[k] = np.float64(self.hits[k]) / np.float64(self.seen)
        return self.metric

    def log_metric(self, epoch: Optional[int] = None, to_tensorboard: bool = True) -> None:
        for k in self.ks:
            logger.info(f"Start of epoch {epoch} recall@{k}: {self.metric[k]}")
            if to_tensorboard and _SUMMARY_WRITER[0] is not None:
                _SUMMARY_WRITER[0].add_scalar(f"Epoch Start Recall@{k}", float(self.metric[k]), epoch)
