"""
StaticIndex: the same fixed, ordered candidate list for every query -- rule-based baselines such as
popularity (reference pkg/modelling/indices/static_index.py:9-95).  No numerics: host-side numpy.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from pkg.modelling import _device as D
from pkg.modelling.models.abstract_keras_model import AbstractKerasModel, TensorSpec
from pkg.schema.features import Feature
from pkg.schema.schema import Schema


class StaticIndex(AbstractKerasModel):
    def __init__(self, k: int, input_features: List[Feature], candidates):
        """``candidates``: ordered ids, shape (1, num_candidates)."""
        super().__init__()
        self.k = int(k)
        self.input_features = input_features
        self.candidates = np.asarray(candidates).reshape(1, -1)
        self.initialise_model()

    def call(self, x, training: bool = False) -> np.ndarray:
        num_results = D.batch_size_of(x[self.input_features[0].name])
        return np.tile(self.candidates[:, : self.k], (num_results, 1))

    def get_input_signature(self) -> Dict[str, TensorSpec]:
        return {f.name: TensorSpec((None, 1), f.dtype, f.name) for f in self.input_features}

    @classmethod
    def build_popularity_index_from_series_schema(cls, schema: Schema, s) -> "StaticIndex":
        """Ids ordered by frequency in the pandas Series ``s`` (value_counts order)."""
        ids = s.value_counts().index
        return cls(k=max(schema.model_config.ks), input_features=schema.query_features,
                   candidates=np.array([str(i) for i in ids]).reshape(1, len(ids)))      # fixed-width strings: cheap to tile and compare

    def state_arrays(self) -> Dict[str, np.ndarray]:
        return {"candidates": self.candidates.astype(str)}
