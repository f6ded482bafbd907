"""
CPU oracle for the two-tower hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product package (``hm-retrieval-two-tower_b200/pkg``) never does: it fails loudly when
the CUDA library is missing.

What is restated (file:line under /root/reference):

* StringLookup(num_oov_indices=1) + Embedding + Concatenate   pkg/modelling/layers/input_layer.py:24-43,61-69
* Dense(relu) stack incl. ReLU on the last layer               pkg/modelling/models/tower.py:41-49,72-75
* logits = q @ c.T                                             pkg/modelling/models/two_tower_model.py:86-92
* logQ correction  z = s - ln p(id_j), default p = 1           pkg/modelling/layers/logq_correction.py:32-42,66-71
* labels = eye(B), CategoricalCrossentropy(from_logits, SUM)   two_tower_model.py:110-124, pkg/modelling/runner.py:78-83
* optimizer.minimize with tf-keras 2.16 legacy Adagrad / Adam  pkg/modelling/optimizer_factory.py:15-18
* brute-force scoring + top_k + id gather                      pkg/modelling/indices/brute_force.py:40-52,75-83,97-106
* StaticIndex tile                                             pkg/modelling/indices/static_index.py:54-55
* Recall@k hit counting                                        pkg/modelling/metrics/index_recall.py:52-58

The arithmetic itself lives in third-party code that is NOT under /root/reference: tensorflow 2.16.2
and tf-keras 2.16.0 (poetry.lock).  Neither is importable in this image, so the reference cannot be
run here; the rules encoded below are the published semantics of those versions.

Pinning status
--------------
PINNED against the reference's own fixtures (tests/test_oracle.py, tests/golden/reference_fixtures.json):
  logQ 3x3 (tests/test_layers.py:8-36), brute-force top-2 with OOV query (tests/test_indices.py:63-129),
  Recall@{1,2,5} over ragged batches (tests/test_recall.py:8-95).
PARITY UNPINNED (no reference test or runnable reference covers them): embedding gather / concat order,
  Dense/ReLU, CE-SUM value, every gradient, Adagrad/Adam (dense and sparse-dedup), top_k tie order,
  StaticIndex.call.  These follow upstream documentation and are frozen by the known-answer vectors
  KAT-A/B/C in tests/golden/kat.json, and cross-checked without TensorFlow in tests/test_oracle.py: the gradients against
  torch autograd of an independently written float64 forward and against finite differences of the loss; sparse Adagrad
  against the dense rule on the densified gradient and torch.optim.Adagrad (the same update rule).

Two evaluation modes are offered for contractions:
  ``canonical=True``  sequential k-ascending fused multiply-add in fp32 (C helper, oracle/tt_oracle.c);
                      this is the order the exact CUDA paths implement, so results compare bit for bit.
  ``canonical=False`` float64 accumulation rounded once to fp32 ("truth" for tolerance checks).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtt_oracle.so")
_lib = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/tt_oracle.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "tt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def c_lib():
    global _lib
    if _lib is None:
        build_c_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        i64, f32p, i64p, i32p = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        lib.tto_gemm_nt_fmaf.argtypes = [f32p, f32p, f32p, i64, i64, i64, i64, i64, i64]
        lib.tto_gemm_nt_fmaf.restype = None
        lib.tto_dense_fmaf.argtypes = [f32p, f32p, f32p, f32p, i64, i64, i64, i64, i64, ctypes.c_int]
        lib.tto_dense_fmaf.restype = None
        lib.tto_topk.argtypes = [f32p, i64, i64, i64, f32p, i64p]
        lib.tto_topk.restype = ctypes.c_int
        lib.tto_index_topk.argtypes = [f32p, f32p, i64, i64, i64, i64, i64, f32p, i64p]
        lib.tto_index_topk.restype = ctypes.c_int
        lib.tto_recall_hits.argtypes = [i64p, i64p, i64, i64, i32p, ctypes.c_int32, i32p]
        lib.tto_recall_hits.restype = None
        lib.tto_max_threads.restype = ctypes.c_int
        lib.tto_set_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def _f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------
# vocabulary / input layer
# --------------------------------------------------------------------------------------------
def string_lookup(vocab: Sequence[str], values) -> np.ndarray:
    """StringLookup(num_oov_indices=1, vocabulary=vocab): OOV -> 0, vocab[i] -> i + 1 (int64).
    input_layer.py:33-36."""
    table = {}
    for i, v in enumerate(vocab):
        table.setdefault(v.decode() if isinstance(v, bytes) else str(v), i + 1)
    flat = np.asarray(values).reshape(-1)
    out = np.empty(flat.shape[0], dtype=np.int64)
    for n, v in enumerate(flat):
        out[n] = table.get(v.decode() if isinstance(v, bytes) else str(v), 0)
    return out.reshape(np.asarray(values).shape)


@dataclass
class OracleFeature:
    name: str
    is_string: bool
    embedding_size: Optional[int] = None


@dataclass
class OracleTower:
    """Parameters of one Tower (tower.py:36-49): tables keyed by feature *name* (the reference keeps a
    dict keyed by name, input_layer.py:30-43, so same-named features share the last-built table),
    then Dense layers as (W (in,out), b (out,))."""
    features: List[OracleFeature]
    tables: Dict[str, np.ndarray]
    dense: List[Tuple[np.ndarray, np.ndarray]]

    @property
    def numerical(self) -> List[OracleFeature]:
        return [f for f in self.features if not f.is_string]

    @property
    def categorical(self) -> List[OracleFeature]:
        return [f for f in self.features if f.is_string]


def input_layer_forward(tower: OracleTower, ids: Dict[str, np.ndarray], numerics: Dict[str, np.ndarray]):
    """concat([numerics in schema order] + [table[name][ids] in schema order], axis=-1).
    input_layer.py:61-69.  ``ids[name]`` are row indices (0 = OOV), shape (B,) or (B,1)."""
    cols = []
    for f in tower.numerical:
        cols.append(np.asarray(numerics[f.name], dtype=np.float32).reshape(-1, 1))
    for f in tower.categorical:
        t = tower.tables[f.name]
        cols.append(t[np.asarray(ids[f.name]).reshape(-1)])
    return np.concatenate(cols, axis=1).astype(np.float32)


def dense_relu(x: np.ndarray, w: np.ndarray, b: np.ndarray, canonical: bool = False) -> np.ndarray:
    """relu(x @ w + b)   (tower.py:44-49; tf-keras Dense, activation="relu")."""
    if canonical:
        x, w, b = _f32c(x), _f32c(w), _f32c(b)
        y = np.empty((x.shape[0], w.shape[1]), dtype=np.float32)
        c_lib().tto_dense_fmaf(_ptr(x), _ptr(w), _ptr(b), _ptr(y), x.shape[0], w.shape[1], x.shape[1],
                               x.shape[1], w.shape[1], 1)
        return y
    y = x.astype(np.float64) @ w.astype(np.float64) + b.astype(np.float64)
    return np.maximum(y, 0.0).astype(np.float32)


def tower_forward(tower: OracleTower, ids, numerics, canonical: bool = False) -> List[np.ndarray]:
    """Returns [x0, h1, ..., hL]; hL is the (B, E) tower output.  tower.py:72-75."""
    acts = [input_layer_forward(tower, ids, numerics)]
    for w, b in tower.dense:
        acts.append(dense_relu(acts[-1], w, b, canonical))
    return acts


# --------------------------------------------------------------------------------------------
# logits, logQ, loss
# --------------------------------------------------------------------------------------------
def logits_qct(q: np.ndarray, c: np.ndarray, canonical: bool = False) -> np.ndarray:
    """q @ c.T  (two_tower_model.py:92, brute_force.py:76-78)."""
    if canonical:
        q, c = _f32c(q), _f32c(c)
        out = np.empty((q.shape[0], c.shape[0]), dtype=np.float32)
        c_lib().tto_gemm_nt_fmaf(_ptr(q), _ptr(c), _ptr(out), q.shape[0], c.shape[0], q.shape[1],
                                 q.shape[1], c.shape[1], c.shape[0])
        return out
    return (q.astype(np.float64) @ c.astype(np.float64).T).astype(np.float32)


def round_tf32(x: np.ndarray) -> np.ndarray:
    """fp32 -> TF32 (10 explicit mantissa bits), round to nearest with ties away from zero: what cvt.rna.tf32.f32 and the
    product's tt_round_tf32 produce for finite values.  Used to restate the operands of the tensor-core index filter."""
    bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def logq_correction(logits: np.ndarray, probs: np.ndarray) -> np.ndarray:
    """logits - log(p)[None, :] in fp32 (logq_correction.py:66-71).  ``probs`` is the per-column
    sampling probability after the table lookup (default 1.0 for unknown ids, :38-41)."""
    corr = np.log(np.asarray(probs, dtype=np.float32)).reshape(1, -1)
    return (np.asarray(logits, dtype=np.float32) - corr).astype(np.float32)


def prob_lookup(table: Dict[str, float], candidate_ids) -> np.ndarray:
    """StaticHashTable[str -> float32], default 1.0 (logq_correction.py:32-42)."""
    flat = np.asarray(candidate_ids).reshape(-1)
    out = np.ones(flat.shape[0], dtype=np.float32)
    for n, v in enumerate(flat):
        key = v.decode() if isinstance(v, bytes) else str(v)
        if key in table:
            out[n] = np.float32(table[key])
    return out


def ce_sum_from_logits(z: np.ndarray, diag_offset: int = 0):
    """labels = eye; CategoricalCrossentropy(from_logits=True, reduction=SUM)
    (two_tower_model.py:119-122, runner.py:78-83).  Returns (loss f64, lse f64 (B,), dZ f64 (B,Bc)).
    Row i's positive is column i + diag_offset (all-gathered negatives)."""
    z64 = np.asarray(z, dtype=np.float64)
    m = z64.max(axis=1, keepdims=True)
    e = np.exp(z64 - m)
    s = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(s)).reshape(-1)
    rows = np.arange(z64.shape[0])
    loss = float(np.sum(lse - z64[rows, rows + diag_offset]))
    dz = e / s
    dz[rows, rows + diag_offset] -= 1.0
    return loss, lse, dz


# --------------------------------------------------------------------------------------------
# full train-step gradients
# --------------------------------------------------------------------------------------------
@dataclass
class IndexedSlices:
    """tf.IndexedSlices as produced by the Embedding gradient: duplicates NOT merged."""
    indices: np.ndarray  # (n,)
    values: np.ndarray   # (n, e) fp32


@dataclass
class StepGrads:
    loss: float
    logits: np.ndarray
    q: np.ndarray
    c: np.ndarray
    dq: np.ndarray
    dc: np.ndarray
    dense_q: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    dense_c: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    tables_q: Dict[str, IndexedSlices] = field(default_factory=dict)
    tables_c: Dict[str, IndexedSlices] = field(default_factory=dict)


def _tower_backward(tower: OracleTower, acts: List[np.ndarray], ids, dout: np.ndarray):
    """Backprop through Dense(relu) stack and the concat/gather (autodiff of tower.py:72-75 and
    input_layer.py:61-69).  fp64 internally, rounded to fp32 at the end."""
    dense_grads = []
    d = dout.astype(np.float64)
    for layer in range(len(tower.dense) - 1, -1, -1):
        w, _ = tower.dense[layer]
        y = acts[layer + 1]
        x = acts[layer].astype(np.float64)
        dpre = d * (y > 0)
        dw = x.T @ dpre
        db = dpre.sum(axis=0)
        d = dpre @ w.astype(np.float64).T
        dense_grads.append((dw.astype(np.float32), db.astype(np.float32)))
    dense_grads.reverse()
    dx = d.astype(np.float32)
    # split the concat: numerics first, then categoricals, schema order
    off = len(tower.numerical)
    slices: Dict[str, IndexedSlices] = {}
    for f in tower.categorical:
        e = tower.tables[f.name].shape[1]
        idx = np.asarray(ids[f.name]).reshape(-1).astype(np.int64)
        val = dx[:, off:off + e]
        if f.name in slices:  # same-named feature sharing one table: slices concatenate
            slices[f.name] = IndexedSlices(np.concatenate([slices[f.name].indices, idx]),
                                           np.concatenate([slices[f.name].values, val], axis=0))
        else:
            slices[f.name] = IndexedSlices(idx, np.ascontiguousarray(val))
        off += e
    return dense_grads, slices, dx


def train_step_grads(qt: OracleTower, ct: OracleTower, q_ids, q_num, c_ids, c_num,
                     col_probs: Optional[np.ndarray]) -> StepGrads:
    """One TwoTowerModel.train_step up to (not including) the optimizer (two_tower_model.py:94-124)."""
    qa = tower_forward(qt, q_ids, q_num)
    ca = tower_forward(ct, c_ids, c_num)
    q, c = qa[-1], ca[-1]
    s = logits_qct(q, c)
    z = logq_correction(s, col_probs) if col_probs is not None else s
    loss, _, dz = ce_sum_from_logits(z)
    dq = dz @ c.astype(np.float64)
    dc = dz.T @ q.astype(np.float64)
    gq, sq, _ = _tower_backward(qt, qa, q_ids, dq)
    gc, sc, _ = _tower_backward(ct, ca, c_ids, dc)
    return StepGrads(loss, z, q, c, dq.astype(np.float32), dc.astype(np.float32), gq, gc, sq, sc)


# --------------------------------------------------------------------------------------------
# optimizers (tf-keras 2.16.0 legacy; optimizer_factory.py:15-18)
# --------------------------------------------------------------------------------------------
ADAGRAD_INIT_ACC = np.float32(0.1)
KERAS_EPS = np.float32(1e-7)


DEDUP_BLOCK = 32


def dedup_indexed_slices(s: IndexedSlices, block: int = DEDUP_BLOCK) -> IndexedSlices:
    """OptimizerV2._deduplicate_indexed_slices: Unique + UnsortedSegmentSum.  TensorFlow leaves the
    summation order of duplicates unspecified; the canonical fp32 order fixed here (and implemented by
    tt_sparse.cu) is: stable-sort the (id, position) pairs by id; cut the sorted array into blocks of
    ``block`` entries; within a block, sum a run of equal ids sequentially in ascending position; add the
    per-block pieces of a run in block order.  Runs shorter than a block that do not straddle a block
    boundary -- the common case -- are plain sequential sums.  Unique ids come back in ascending order."""
    idx = np.asarray(s.indices, dtype=np.int64)
    vals = s.values.astype(np.float32)
    order = np.argsort(idx, kind="stable")
    sid = idx[order]
    uniq, starts = np.unique(sid, return_index=True)
    ends = np.append(starts[1:], sid.shape[0])
    out = np.zeros((uniq.shape[0], vals.shape[1]), dtype=np.float32)
    for u, (a, b) in enumerate(zip(starts, ends)):
        total = None
        lo = a
        while lo < b:
            hi = min(b, (lo // block + 1) * block)
            piece = np.zeros(vals.shape[1], dtype=np.float32)
            for j in range(lo, hi):                      # sequential fp32 adds, ascending position
                piece = piece + vals[order[j]]
            total = piece if total is None else total + piece
            lo = hi
        out[u] = total
    return IndexedSlices(uniq, out)


def adagrad_dense(w: np.ndarray, acc: np.ndarray, g: np.ndarray, lr: float, eps=KERAS_EPS):
    """ResourceApplyAdagradV2: acc += g*g ; w -= (g*lr) / (sqrt(acc) + eps).  All fp32, every op
    rounded separately (the CUDA kernel uses __fmul_rn/__fadd_rn/__fsqrt_rn/__fdiv_rn to match)."""
    g = g.astype(np.float32)
    lr32 = np.float32(lr)
    acc += g * g
    w -= (g * lr32) / (np.sqrt(acc) + np.float32(eps))


def adagrad_sparse(table: np.ndarray, acc: np.ndarray, s: IndexedSlices, lr: float, eps=KERAS_EPS):
    """Duplicate ids are summed first, then the Adagrad formula on the touched rows only."""
    d = dedup_indexed_slices(s)
    g = d.values
    a = acc[d.indices] + g * g
    acc[d.indices] = a
    table[d.indices] = table[d.indices] - (g * np.float32(lr)) / (np.sqrt(a) + np.float32(eps))


def adam_lr_t(lr: float, step: int, b1=0.9, b2=0.999) -> np.float32:
    return np.float32(lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step))


def adam_dense(w, m, v, g, lr, step, b1=0.9, b2=0.999, eps=KERAS_EPS):
    """ResourceApplyAdam (legacy Adam, amsgrad off): m += (g-m)(1-b1); v += (g*g-v)(1-b2);
    w -= lr_t*m/(sqrt(v)+eps)."""
    g = g.astype(np.float32)
    a = adam_lr_t(lr, step, b1, b2)
    # tf-keras forms (1 - beta) from the fp32 hyper-parameter tensors: fp32 subtraction, not float64
    m += (g - m) * (np.float32(1.0) - np.float32(b1))
    v += (g * g - v) * (np.float32(1.0) - np.float32(b2))
    w -= (m * a) / (np.sqrt(v) + np.float32(eps))


def adam_sparse(table, m, v, s: IndexedSlices, lr, step, b1=0.9, b2=0.999, eps=KERAS_EPS):
    """legacy Adam._resource_apply_sparse is NOT lazy: m and v decay over the whole table, the
    de-duplicated gradient is scatter-added, and every row's weight moves."""
    d = dedup_indexed_slices(s)
    a = adam_lr_t(lr, step, b1, b2)
    m *= np.float32(b1)
    m[d.indices] += d.values * (np.float32(1.0) - np.float32(b1))
    v *= np.float32(b2)
    v[d.indices] += (d.values * d.values) * (np.float32(1.0) - np.float32(b2))
    table -= (m * a) / (np.sqrt(v) + np.float32(eps))


# --------------------------------------------------------------------------------------------
# index, top-k, recall
# --------------------------------------------------------------------------------------------
def top_k(scores: np.ndarray, k: int):
    """tf.math.top_k: sorted descending; equal values -> lower index first (brute_force.py:81)."""
    s = _f32c(scores)
    nq, n = s.shape
    out_s = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    rc = c_lib().tto_topk(_ptr(s), nq, n, k, _ptr(out_s), _ptr(out_i))
    assert rc == 0
    return out_s, out_i


def top_k_numpy(scores: np.ndarray, k: int):
    """Independent pure-numpy statement of the same rule (stable argsort of -score)."""
    s = np.asarray(scores, dtype=np.float32)
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, axis=1), order.astype(np.int64)


def index_topk(q: np.ndarray, corpus: np.ndarray, k: int, idx_base: int = 0, threads: int = 0):
    """BruteForceIndex.call after the query tower: canonical scores + top-k, fused (C, OpenMP)."""
    q, corpus = _f32c(q), _f32c(corpus)
    lib = c_lib()
    if threads:
        lib.tto_set_threads(threads)
    out_s = np.empty((q.shape[0], k), dtype=np.float32)
    out_i = np.empty((q.shape[0], k), dtype=np.int64)
    rc = lib.tto_index_topk(_ptr(q), _ptr(corpus), q.shape[0], corpus.shape[0], q.shape[1], k, idx_base,
                            _ptr(out_s), _ptr(out_i))
    assert rc == 0
    return out_s, out_i


def merge_topk(scores: np.ndarray, idx: np.ndarray, k: int):
    """K-way merge of per-shard results (G, nq, k) by (score desc, idx asc)."""
    g, nq, kk = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(nq, g * kk)
    i = np.transpose(idx, (1, 0, 2)).reshape(nq, g * kk)
    order = np.lexsort((i, -s.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)


def static_index_call(candidates: np.ndarray, k: int, batch: int) -> np.ndarray:
    """tile(candidates[:, :k], (B, 1))   static_index.py:54-55."""
    return np.tile(np.asarray(candidates)[:, :k], (batch, 1))


class RecallOracle:
    """index_recall.py:22-59: int32 hit / seen counters, float64 ratio."""

    def __init__(self, ks: Sequence[int]):
        self.ks = list(ks)
        self.hits = {k: np.int32(0) for k in self.ks}
        self.seen = np.int32(0)
        self.metric = {k: np.int32(0) for k in self.ks}

    def update(self, true_ids: np.ndarray, candidates: np.ndarray):
        t = np.asarray(true_ids).reshape(-1, 1)
        self.seen = np.int32(self.seen + t.shape[0])
        for k in self.ks:
            eq = (t == np.asarray(candidates)[:, :k])
            self.hits[k] = np.int32(self.hits[k] + np.int32(eq.sum()))
            self.metric[k] = np.float64(self.hits[k]) / np.float64(self.seen)
        return self.metric
