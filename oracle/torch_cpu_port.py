"""
Timed CPU baseline ("port") -- TEST / BENCH INFRASTRUCTURE ONLY, never on the product path.

The reference's hot path executed with torch-CPU ops (MKL / oneDNN sgemm, the same class of kernels
TensorFlow 2.16 dispatches to on x86), multi-threaded, with the input pipeline and string lookups
excluded exactly as on the GPU side.  TensorFlow itself is not installable in this image, so this is a
restatement and every report labels it "port".

Follows (file:line under /root/reference): two_tower_model.py:94-130 (train_step), input_layer.py:61-69,
tower.py:72-75, logq_correction.py:66-71, runner.py:78-83 (CE from logits, SUM), optimizer_factory.py:15-18
(legacy Adagrad, sparse path de-duplicated), brute_force.py:75-83 (matmul + top_k).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


class CpuTower:
    def __init__(self, cat_sizes: List[Tuple[int, int]], n_numeric: int, units: List[int], gen: torch.Generator):
        """cat_sizes: [(rows, e)] in schema order; numerics come first in the concat."""
        self.tables = [torch.empty(r, e).uniform_(-0.05, 0.05, generator=gen) for r, e in cat_sizes]
        self.accs = [torch.full_like(t, 0.1) for t in self.tables]
        d = n_numeric + sum(e for _, e in cat_sizes)
        self.dense = []
        for n in units:
            lim = (6.0 / (d + n)) ** 0.5
            w = torch.empty(d, n).uniform_(-lim, lim, generator=gen)
            b = torch.zeros(n)
            self.dense.append([w, b, torch.full_like(w, 0.1), torch.full_like(b, 0.1)])
            d = n
        self.n_numeric = n_numeric

    def forward(self, ids: List[torch.Tensor], nums: List[torch.Tensor]):
        x = torch.cat([n.reshape(-1, 1) for n in nums] + [t.index_select(0, i) for t, i in zip(self.tables, ids)], dim=1)
        acts = [x]
        for w, b, _, _ in self.dense:
            acts.append(torch.relu(torch.addmm(b, acts[-1], w)))
        return acts

    def backward_and_update(self, acts, ids, d_out, lr, eps=1e-7):
        d = d_out
        for li in range(len(self.dense) - 1, -1, -1):
            w, b, aw, ab = self.dense[li]
            dpre = d * (acts[li + 1] > 0)
            dw = acts[li].t() @ dpre
            db = dpre.sum(0)
            d = dpre @ w.t()
            aw.addcmul_(dw, dw); w.addcdiv_(dw, aw.sqrt() + eps, value=-lr)
            ab.addcmul_(db, db); b.addcdiv_(db, ab.sqrt() + eps, value=-lr)
        off = self.n_numeric
        for t, acc, i in zip(self.tables, self.accs, ids):
            e = t.shape[1]
            g = d[:, off:off + e]
            off += e
            uniq, inv = torch.unique(i, return_inverse=True)          # Unique
            gs = torch.zeros(uniq.shape[0], e).index_add_(0, inv, g)  # UnsortedSegmentSum
            a = acc.index_select(0, uniq) + gs * gs                   # ResourceSparseApplyAdagradV2
            acc.index_copy_(0, uniq, a)
            t.index_copy_(0, uniq, t.index_select(0, uniq) - lr * gs / (a.sqrt() + eps))


class CpuTwoTower:
    def __init__(self, q_cat, q_num, c_cat, c_num, joint, q_units=None, c_units=None, log_p_rows: Optional[torch.Tensor] = None,
                 lr: float = 0.05, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.q = CpuTower(q_cat, q_num, list(q_units or []) + [joint], g)
        self.c = CpuTower(c_cat, c_num, list(c_units or []) + [joint], g)
        self.log_p_rows = log_p_rows
        self.lr = lr

    def train_step(self, q_ids, q_nums, c_ids, c_nums) -> float:
        qa = self.q.forward(q_ids, q_nums)
        ca = self.c.forward(c_ids, c_nums)
        s = qa[-1] @ ca[-1].t()
        if self.log_p_rows is not None:
            s = s - self.log_p_rows.index_select(0, c_ids[0]).reshape(1, -1)
        lsm = torch.log_softmax(s, dim=1)
        loss = -lsm.diagonal().sum()
        dz = lsm.exp_()
        dz.diagonal().sub_(1.0)
        dq = dz @ ca[-1]
        dc = dz.t() @ qa[-1]
        self.q.backward_and_update(qa, q_ids, dq, self.lr)
        self.c.backward_and_update(ca, c_ids, dc, self.lr)
        return float(loss)


def cpu_index_topk(q: torch.Tensor, corpus: torch.Tensor, k: int):
    """brute_force.py:75-83 on the CPU: full (B, N) score matrix, then top_k."""
    scores = q @ corpus.t()
    return torch.topk(scores, k, dim=1, largest=True, sorted=True)


def time_fn(fn, warmup: int, steps: int) -> float:
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps
