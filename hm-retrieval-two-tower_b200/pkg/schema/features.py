"""
Feature metadata feeding the hot path: vocab order fixes the embedding row of every id
(row 0 = OOV, vocab[i] -> row i+1), mirroring reference pkg/schema/features.py:11-127 without TensorFlow.
"""
from __future__ import annotations

import logging
from enum import Enum
from typing import List, Optional, Sequence

import numpy as np

from pkg.schema.dtypes import DType, as_dtype

logger = logging.getLogger(__name__)


class FeatureFamily(Enum):
    QUERY = "query"
    CANDIDATE = "candidate"


class Feature:
    """One model input (reference features.py:21-81: same arguments, same TypeError / ValueError cases).

    ``vocab`` keeps first-occurrence order (the reference stores ``set(vocab)``, whose iteration order is
    arbitrary; a deterministic order is needed for reproducible row ids).
    """

    VALID_DTYPES = [DType.string, DType.float32]

    def __init__(self, name: str, dtype, feature_family: FeatureFamily, embedding_size: Optional[int] = None,
                 vocab: Optional[Sequence[str]] = None, max_vocab_size: Optional[int] = None):
        self.name = name
        self.dtype = as_dtype(dtype)  # TypeError for anything but string / float32
        if not isinstance(feature_family, FeatureFamily):
            raise ValueError(f"feature_family {feature_family} not valid. Must be one of {FeatureFamily._member_names_}")
        self.feature_family = feature_family
        if embedding_size and self.dtype != DType.string:
            raise TypeError(f"Got embedding size, dtype must be string got {self.dtype}")
        self.embedding_size = embedding_size
        if self.dtype != DType.string:
            if vocab:
                logger.info(f"Ignoring vocab passed for non-string feature {self.name}")
            self.vocab, self.is_built = None, True
        elif vocab is not None and len(vocab) > 0:
            self.vocab, self.is_built = list(dict.fromkeys(str(v) for v in vocab)), True
        else:
            self.vocab, self.is_built = None, False
        if max_vocab_size and not isinstance(max_vocab_size, int):
            raise TypeError(f"max_vocab_size must be an int, got {max_vocab_size}")
        self.max_vocab_size = max_vocab_size

    def set_vocab_from_dataframe(self, df) -> None:
        """Most frequent value first (``value_counts`` order), truncated to ``max_vocab_size``
        (reference features.py:106-127)."""
        if self.name not in df.columns:
            raise ValueError(f"Feature name {self.name} not found in df cols {df.columns}")
        counts = df[self.name].value_counts()
        keep = counts.head(self.max_vocab_size) if self.max_vocab_size else counts
        self.vocab = np.array([str(v) for v in keep.index])
        self.is_built = True

    def set_vocab_size(self, n: int) -> None:
        """Synthetic vocabulary "1".."n" for pre-encoded integer ids (row id == int(value))."""
        self.vocab = _RangeVocab(n)
        self.is_built = True

    def __repr__(self) -> str:
        return f"Feature({self.name!r}, {self.dtype!r}, {self.feature_family.name}, e={self.embedding_size})"


class _RangeVocab:
    """Lazy vocabulary ["1", ..., "n"]: avoids materialising 10^8 Python strings for synthetic tables."""

    def __init__(self, n: int):
        self.n = int(n)

    def __len__(self) -> int:
        return self.n

    def __iter__(self):
        return (str(i + 1) for i in range(self.n))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [str(j + 1) for j in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return str(i + 1)
