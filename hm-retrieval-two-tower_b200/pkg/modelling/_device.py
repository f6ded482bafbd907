"""
Host<->device plumbing shared by the pkg.modelling classes: vocabulary lookup (the StringLookup half of
reference input_layer.py:33-36, done on the host because strings never reach the GPU), staging of a
batch into device buffers, and a tiny parameter store that keeps all Dense weights of a model in one
flat buffer (one optimizer launch, one all-reduce).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from pkg import _native as N


def _as_str(v) -> str:
    return v.decode() if isinstance(v, (bytes, np.bytes_)) else str(v)


def _as_bytes_array(values) -> np.ndarray:
    """Any array of str / bytes / objects -> contiguous numpy 'S<w>' array of shape (n,)."""
    flat = np.asarray(values).reshape(-1)
    if flat.dtype.kind == "S":
        return np.ascontiguousarray(flat)
    if flat.dtype.kind == "U":
        try:
            return np.ascontiguousarray(flat.astype("S"))          # ASCII fast path
        except UnicodeEncodeError:
            return np.ascontiguousarray(np.char.encode(flat, "utf-8"))
    return np.ascontiguousarray(np.array([v if isinstance(v, (bytes, np.bytes_)) else str(v).encode() for v in flat], dtype="S"))


class Vocab:
    """StringLookup(num_oov_indices=1, vocabulary=vocab): OOV -> 0, vocab[i] -> i + 1.

    Lookups of string batches run in libtt's open-addressing hash map (tt_vocab_lookup_fixed, multi-threaded): the
    per-example Python dict lookup this replaces was the first host bottleneck once the kernels were fast."""

    def __init__(self, vocab):
        self.size = len(vocab)
        self._range = hasattr(vocab, "n") and vocab.__class__.__name__ == "_RangeVocab"
        self._handle = None
        self._vocab = vocab

    @property
    def rows(self) -> int:
        return self.size + 1

    def _native(self):
        if self._handle is None:
            lib = N.load()
            enc = [(_as_str(v)).encode() for v in self._vocab]
            offsets = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(b) for b in enc], out=offsets[1:])
            blob = b"".join(enc)
            self._handle = lib.tt_vocab_create(blob, offsets.ctypes.data, len(enc))
            if not self._handle:
                raise N.TTError("tt_vocab_create failed: " + lib.tt_last_error().decode())
        return self._handle

    def __del__(self):
        try:
            if self._handle:
                N.load().tt_vocab_destroy(self._handle)
        except Exception:
            pass

    def __getstate__(self):          # the native handle does not pickle; it is rebuilt on demand
        d = dict(self.__dict__)
        d["_handle"] = None
        return d

    def encode(self, values, nthreads: int = 8) -> np.ndarray:
        """Any array of str / bytes / objects -> int32 row ids, shape (n,)."""
        cells = _as_bytes_array(values)
        n = cells.shape[0]
        out = np.zeros(n, dtype=np.int32)
        if n == 0:
            return out
        if self._range:   # synthetic vocabulary "1".."n": the row id is the integer itself when it is canonical and in range
            for i, v in enumerate(cells):
                s = v.decode()
                if s.isdigit() and str(int(s)) == s and 1 <= int(s) <= self.size:
                    out[i] = int(s)
            return out
        lib = N.load()
        N.check(lib.tt_vocab_lookup_fixed(self._native(), cells.ctypes.data, n, cells.dtype.itemsize, out.ctypes.data, nthreads),
                "tt_vocab_lookup_fixed")
        return out

    def token(self, row: int) -> str:
        return "[UNK]" if row == 0 else _as_str(self._vocab[row - 1])


def unwrap(value):
    """Foreign tensors -> torch tensor or numpy array, zero-copy where the producer allows it.

    The reference feeds its layers TF/Keras tensors (input_layer.py:45-69).  Numeric TF EagerTensors, CuPy and JAX arrays
    export ``__dlpack__``: they are imported with ``torch.from_dlpack`` (a device tensor stays on the device and is read in
    place by the staging copy).  TF string tensors have no DLPack form; they come through ``.numpy()`` (an object array of
    bytes) and take the host vocabulary lookup like any other string batch."""
    import torch
    if isinstance(value, (torch.Tensor, np.ndarray, list, tuple)) or np.isscalar(value):
        return value
    if hasattr(value, "__dlpack__"):
        try:
            return torch.from_dlpack(value)
        except Exception:      # string / ragged / unsupported dtype: fall through to the host view
            pass
    if hasattr(value, "numpy"):
        return value.numpy()
    return value


def is_string_like(a) -> bool:
    a = unwrap(a)
    if isinstance(a, np.ndarray):
        return a.dtype.kind in ("U", "S", "O")
    if isinstance(a, (list, tuple)):
        return len(a) > 0 and isinstance(np.asarray(a).reshape(-1)[0], (str, bytes, np.str_, np.bytes_))
    return False


def batch_size_of(x) -> int:
    shp = getattr(x, "shape", None)
    if shp is None:
        return len(x)
    return int(shp[0]) if len(shp) else 1


class PinRing:
    """A few pinned host blocks of (n_cols, rows) 4-byte cells that host-resident feature columns are packed into before the staging
    kernel reads them in place.  A block is reused only after the kernel that read it has run (one event per block), so the host
    packs batch k+1 while the GPU still works on batch k."""

    def __init__(self, rows: int, n_cols: int, depth: int = 3):
        torch = N.require_cuda()
        self.rows, self.n_cols = int(rows), int(n_cols)
        self.blocks = [torch.empty((max(n_cols, 1), max(rows, 1)), dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.views_i = [b.numpy() for b in self.blocks]
        self.views_f = [v.view(np.float32) for v in self.views_i]
        self.events = [None] * depth
        self.cur = -1

    def next_block(self) -> int:
        self.cur = (self.cur + 1) % len(self.blocks)
        ev = self.events[self.cur]
        if ev is not None:
            ev.synchronize()
        return self.cur

    def mark_in_flight(self, k: int) -> None:
        torch = N.require_cuda()
        if self.events[k] is None:
            self.events[k] = torch.cuda.Event()
        self.events[k].record()


class Stager:
    """Collects the feature columns of one batch and stages them with ONE kernel launch (tt_stage_columns): device tensors and
    pinned host tensors are read where they are, anything else on the host (numpy arrays, strings after the vocabulary lookup)
    is packed into a PinRing block first."""

    def __init__(self, rows: int, ring_owner):
        self.rows = int(rows)
        self.owner = ring_owner            # any object; the ring is cached on it as ``_pin_ring``
        self.cols = []                     # (src pointer, dst pointer, kind)
        self.keep = []
        self.host = []                     # (numpy array, dst tensor, is_float)

    def _tensor(self, t, out, is_float: bool) -> bool:
        torch = N.require_cuda()
        t = t.reshape(-1)
        if t.numel() != self.rows or not t.is_contiguous():
            return False
        if not (t.is_cuda or t.is_pinned()):
            return False
        if t.is_cuda and t.device != out.device:
            return False
        want = torch.float32 if is_float else torch.int32
        if t.dtype == want:
            kind = 0
        elif not is_float and t.dtype == torch.int64:
            kind = 1
        else:
            return False
        self.cols.append((t.data_ptr(), out.data_ptr(), kind))
        self.keep.append(t)
        return True

    def add_ids(self, value, vocab: "Vocab", out) -> None:
        torch = N.require_cuda()
        value = unwrap(value)
        if isinstance(value, torch.Tensor):
            if not self._tensor(value, out, False):
                out.copy_(value.reshape(-1), non_blocking=True)
            return
        if is_string_like(value):
            ids = vocab.encode(value)
        else:
            ids = np.asarray(value).reshape(-1)
        self.host.append((ids, out, False))

    def add_floats(self, value, out) -> None:
        torch = N.require_cuda()
        value = unwrap(value)
        if isinstance(value, torch.Tensor):
            if not self._tensor(value, out, True):
                out.copy_(value.reshape(-1), non_blocking=True)
            return
        self.host.append((np.asarray(value).reshape(-1), out, True))

    def flush(self) -> None:
        lib = N.load()
        ring, blk = None, -1
        if self.host:
            ring = getattr(self.owner, "_pin_ring", None)
            if ring is None or ring.rows != self.rows or ring.n_cols < len(self.host):
                ring = PinRing(self.rows, max(len(self.host), 8))
                self.owner._pin_ring = ring
            blk = ring.next_block()
            base = ring.blocks[blk].data_ptr()
            for j, (arr, out, is_float) in enumerate(self.host):
                if arr.shape[0] != self.rows:
                    raise ValueError(f"feature column of {arr.shape[0]} rows in a batch of {self.rows}")
                dst = (ring.views_f if is_float else ring.views_i)[blk][j]
                np.copyto(dst, arr, casting="unsafe")
                self.cols.append((base + 4 * j * ring.rows, out.data_ptr(), 0))
        n = len(self.cols)
        for lo in range(0, n, N.TT_MAX_STAGE_COLS):
            part = self.cols[lo:lo + N.TT_MAX_STAGE_COLS]
            arr = (N.TTStageCol * len(part))()
            for i, (src, dst, kind) in enumerate(part):
                arr[i].src, arr[i].dst, arr[i].kind = src, dst, kind
            N.check(lib.tt_stage_columns(arr, len(part), self.rows, N.stream_ptr()), "tt_stage_columns")
        if ring is not None:
            ring.mark_in_flight(blk)
        self.cols, self.keep, self.host = [], [], []


class ParamStore:
    """Flat fp32 buffers for Dense kernels/biases: ``params`` and ``grads`` share one layout."""

    def __init__(self, capacity: int):
        torch = N.require_cuda()
        self.capacity = int(max(capacity, 1))
        self.params = torch.zeros(self.capacity, dtype=torch.float32, device="cuda")
        self.grads = torch.zeros(self.capacity, dtype=torch.float32, device="cuda")
        self.used = 0

    def alloc(self, shape: Sequence[int]):
        n = int(np.prod(shape))
        start = (self.used + 3) // 4 * 4  # keep every tensor 16-byte aligned
        if start + n > self.capacity:
            raise RuntimeError("ParamStore overflow")
        self.used = start + n
        return self.params[start:start + n].view(*shape), self.grads[start:start + n].view(*shape)

    @staticmethod
    def padded(n: int) -> int:
        return (n + 3) // 4 * 4


_SEED = [1234]


def set_seed(seed: int) -> None:
    """Seed for parameter initialisation (the reference relies on TF's global seed)."""
    _SEED[0] = int(seed)


def next_seed() -> int:
    """Seed of the next parameter tensor (same sequence as next_generator)."""
    s = _SEED[0]
    _SEED[0] += 1
    return s


def next_generator():
    torch = N.require_cuda()
    g = torch.Generator(device="cuda")
    g.manual_seed(_SEED[0])
    _SEED[0] += 1
    return g


def feature_array(entries: Iterable[dict]):
    """Build a ctypes array of tt_feature from dicts {table, src, rows, e, col} (device pointers)."""
    entries = list(entries)
    arr = (N.TTFeature * len(entries))()
    for i, d in enumerate(entries):
        arr[i].table = d["table"]
        arr[i].src = d["src"]
        arr[i].rows = d["rows"]
        arr[i].e = d["e"]
        arr[i].col = d["col"]
        arr[i].shards = d.get("shards", 0)
    return arr
