"""
TF-free base class keeping the method names of reference pkg/modelling/models/abstract_keras_model.py:10-131
(get_input_signature / set_input_signature / get_default_inputs / initialise_model / call / save).
tf.function tracing and SavedModel export have no meaning without TensorFlow: the signature is kept as
plain metadata and ``save`` writes arrays ("out of scope").
"""
from __future__ import annotations

import logging
import os
from abc import ABC, abstractmethod
from typing import Dict, NamedTuple, Optional, Tuple

import numpy as np

from pkg.schema.dtypes import DType

logger = logging.getLogger(__name__)


class TensorSpec(NamedTuple):
    """Stand-in for tf.TensorSpec: every model input is a (None, 1) column keyed by feature name."""
    shape: Tuple[Optional[int], int]
    dtype: DType
    name: str


class AbstractKerasModel(ABC):
    def __init__(self):
        self._input_signature: Optional[Dict[str, TensorSpec]] = None

    @abstractmethod
    def get_input_signature(self) -> Dict[str, TensorSpec]:
        """Feature name -> TensorSpec((None, 1), dtype, name)."""

    def set_input_signature(self, input_signature: Dict[str, TensorSpec]) -> None:
        self._input_signature = dict(input_signature)

    @staticmethod
    def _get_default_tensor(dtype) -> np.ndarray:
        if dtype == DType.string:
            return np.array([["a"]], dtype=object)
        if dtype == DType.float32:
            return np.zeros((1, 1), dtype=np.float32)
        raise TypeError(f"Invalid dtype {dtype}")

    def get_default_inputs(self, input_signature: Dict[str, TensorSpec]) -> Dict[str, np.ndarray]:
        return {name: self._get_default_tensor(spec.dtype) for name, spec in input_signature.items()}

    @abstractmethod
    def call(self, x, training: bool = True):
        """Run the model on a dict of (B, 1) columns."""

    def __call__(self, x, *args, **kwargs):
        return self.call(x, *args, **kwargs)

    def initialise_model(self) -> None:
        """The reference traces ``call`` on dummy (1,1) inputs (abstract_keras_model.py:109-118); here the
        signature is recorded and its dtypes validated (TypeError on anything but string / float32)."""
        sig = self.get_input_signature()
        self.set_input_signature(sig)
        self.get_default_inputs(sig)

    def state_arrays(self) -> Dict[str, np.ndarray]:
        """Arrays written by ``save`` (overridden by concrete models)."""
        return {}

    def save(self, model_path: str) -> None:
        os.makedirs(os.path.dirname(model_path) or ".", exist_ok=True)
        logger.info(f"Saving model at path: {model_path}")
        os.makedirs(model_path, exist_ok=True)
        np.savez(os.path.join(model_path, "variables.npz"), **self.state_arrays())
