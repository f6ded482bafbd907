"""TFRecord datasets without TensorFlow (reference pkg/modelling/tfrecord_dataset.py:11-98).

``TFRecordDatasetFactory(features).create_tfrecord_dataset(dir, batch_size, shuffle_size)`` keeps the reference's
signature and yields what ``tf.data.TFRecordDataset(...).map(parse).shuffle(s).batch(b)`` yields: dicts
{feature name: (B, 1) array} -- bytes ('S') for string features, float32 for numeric ones; without ``batch_size`` the
elements are single examples of shape (1,).  Files are read whole, their framing and CRCs checked and all Example
messages parsed in one native call (libtt: tt_tfrecord_scan, tt_example_parse); the shuffle is tf.data's buffer shuffle
(a buffer of ``shuffle_size`` elements, a uniformly random one is emitted and replaced by the next input element).
"""
from __future__ import annotations

import ctypes
import logging
import os
from typing import Callable, Dict, Iterator, List, Optional

import numpy as np

from pkg import _native as N
from pkg.schema import dtypes as tt
from pkg.schema.features import Feature

logger = logging.getLogger(__name__)


def read_tfrecord_file(path: str, features: List[Feature], verify_crc: bool = True, nthreads: int = 8) -> Dict[str, np.ndarray]:
    """All examples of one .tfrecord file as columns {name: (n,) array}."""
    lib = N.load()
    data = np.fromfile(path, dtype=np.uint8)
    base = data.ctypes.data
    n = int(lib.tt_tfrecord_scan(base, data.size, 1 if verify_crc else 0, None, None, 0))
    if n < 0:
        raise ValueError(f"{path}: {lib.tt_last_error().decode()}")
    off, ln = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
    if n:
        lib.tt_tfrecord_scan(base, data.size, 0, off.ctypes.data, ln.ctypes.data, n)
    nf = len(features)
    names = (ctypes.c_char_p * nf)(*[f.name.encode() for f in features])
    kind = np.array([0 if f.dtype == tt.string else 1 for f in features], dtype=np.int32)
    s_off, s_len = np.zeros((nf, n), dtype=np.int64), np.zeros((nf, n), dtype=np.int64)
    fv = np.zeros((nf, n), dtype=np.float32)
    if n:
        rc = lib.tt_example_parse(base, off.ctypes.data, ln.ctypes.data, n, names, kind.ctypes.data, nf, s_off.ctypes.data, s_len.ctypes.data,
                                  fv.ctypes.data, nthreads)
        if rc != 0:
            raise ValueError(f"{path}: {lib.tt_last_error().decode()}")
    cols: Dict[str, np.ndarray] = {}
    for j, f in enumerate(features):
        if kind[j] == 1:
            cols[f.name] = fv[j].copy()
            continue
        width = max(int(s_len[j].max()) if n else 1, 1)
        out = np.empty((n, width), dtype=np.uint8)
        # the variable-length byte strings into fixed-width, NUL-padded cells: one native pass (a numpy 'S<width>' column)
        N.check(lib.tt_gather_cells(base, data.size, s_off[j].ctypes.data, s_len[j].ctypes.data, n, width, out.ctypes.data, nthreads), "tt_gather_cells")
        cols[f.name] = out.view(f"S{width}").reshape(n)
    return cols


class TFRecordDataset:
    """Re-iterable stand-in for the tf.data pipeline the reference builds."""

    def __init__(self, filenames: List[str], features: List[Feature], batch_size: Optional[int], shuffle_size: Optional[int], seed: Optional[int] = None,
                 fn: Optional[Callable] = None):
        self.filenames, self.features = filenames, features
        self.batch_size, self.shuffle_size, self.seed, self.fn = batch_size, shuffle_size, seed, fn
        self._cols: Optional[Dict[str, np.ndarray]] = None

    def _load(self) -> Dict[str, np.ndarray]:
        if self._cols is None:
            parts = [read_tfrecord_file(p, self.features) for p in self.filenames]
            self._cols = {f.name: (np.concatenate([p[f.name] for p in parts]) if parts else np.zeros(0, dtype=np.float32)) for f in self.features}
            # string columns of different files may have different widths: numpy widens on concatenate
        return self._cols

    def __len__(self) -> int:
        cols = self._load()
        n = len(next(iter(cols.values()))) if cols else 0
        return n if not self.batch_size else (n + self.batch_size - 1) // self.batch_size

    def _order(self, n: int) -> np.ndarray:
        """Element order of tf.data's buffer shuffle: a buffer of ``shuffle_size`` slots; every step a uniformly random slot emits its
        element and takes the next input element; once the input is exhausted the remaining slots are emitted in random order.
        Vectorised: with r[t] the slot picked at step t, the element a slot emits at one of its picks is the one it took at its
        previous pick (its initial element at the first), so grouping the steps by slot gives the whole order without a Python loop."""
        if not self.shuffle_size or n == 0:
            return np.arange(n)
        rng = np.random.default_rng(self.seed)
        s = min(self.shuffle_size, n)
        m = n - s                                          # steps during which the input still has elements: slot r[t] takes element s + t
        out = np.empty(n, dtype=np.int64)
        content = np.arange(s, dtype=np.int64)             # what the slots hold when the input runs out
        if m:
            r = rng.integers(0, s, size=m)
            by_slot = np.argsort(r, kind="stable")         # steps grouped by slot, ascending within a slot
            rs = r[by_slot]
            new_slot = np.ones(m, dtype=bool)
            new_slot[1:] = rs[1:] != rs[:-1]
            prev_step = np.zeros(m, dtype=np.int64)
            prev_step[1:] = by_slot[:-1]
            out[by_slot] = np.where(new_slot, rs, s + prev_step)
            last_of_slot = np.ones(m, dtype=bool)
            last_of_slot[:-1] = new_slot[1:]
            content[rs[last_of_slot]] = s + by_slot[last_of_slot]
        out[m:] = rng.permutation(content)
        return out

    def map(self, fn: Callable) -> "TFRecordDataset":
        prev = self.fn
        ds = TFRecordDataset(self.filenames, self.features, self.batch_size, self.shuffle_size, self.seed,
                             fn if prev is None else (lambda x: fn(prev(x))))
        ds._cols = self._cols
        return ds

    def __iter__(self) -> Iterator:
        cols = self._load()
        n = len(next(iter(cols.values()))) if cols else 0
        order = self._order(n)
        step = self.batch_size or 1
        for lo in range(0, n, step):
            sel = order[lo:lo + step]
            if self.batch_size:
                item = {k: v[sel].reshape(-1, 1) for k, v in cols.items()}
            else:
                item = {k: v[sel].reshape(1) for k, v in cols.items()}
            yield self.fn(item) if self.fn else item


class TFRecordDatasetFactory:
    def __init__(self, features: List[Feature]):
        self.features = features
        self.feature_description = self._create_feature_description()

    def _create_feature_description(self) -> Dict[str, tuple]:
        """{name: ((1,), dtype)} -- the FixedLenFeature([1], dtype) description of the reference (:22-36)."""
        return {feature.name: ((1,), feature.dtype) for feature in self.features}

    def create_tfrecord_dataset(self, file_dir: str, batch_size: Optional[int] = None, shuffle_size: Optional[int] = None) -> TFRecordDataset:
        filenames = sorted(os.path.join(file_dir, f) for f in os.listdir(file_dir) if f.endswith(".tfrecord"))
        if shuffle_size:
            logger.info(f"Shuffling dataset using shuffle size: {shuffle_size}")
        if batch_size:
            logger.info(f"Batching data using batch_size: {batch_size}")
        return TFRecordDataset(filenames, self.features, batch_size, shuffle_size)
