"""
OptimizerFactory with the reference's contract (pkg/modelling/optimizer_factory.py:15-57): names "adam" /
"adagrad", ``learning_rate`` required, ValueError otherwise.  The optimizers restate tf-keras 2.16.0
``optimizers.legacy`` semantics (dense: ResourceApplyAdagradV2 / ResourceApplyAdam; sparse: duplicates
summed first; Adam's sparse path is not lazy) and run as CUDA kernels (tt_dense_*, tt_sparse_*).
"""
from __future__ import annotations

import logging
import math
from typing import Any, Dict

logger = logging.getLogger(__name__)


class _Optimizer:
    name = "optimizer"

    def __init__(self, learning_rate: float):
        self.learning_rate = float(learning_rate)
        self.iterations = 0

    def get_config(self) -> Dict[str, Any]:
        return {"name": self.name, "learning_rate": self.learning_rate}


class Adagrad(_Optimizer):
    """acc0 = 0.1, eps = 1e-7; acc += g*g; w -= lr*g/(sqrt(acc)+eps)."""
    name = "adagrad"
    n_slots = 1

    def __init__(self, learning_rate: float = 0.001, initial_accumulator_value: float = 0.1, epsilon: float = 1e-7, **_):
        super().__init__(learning_rate)
        if initial_accumulator_value < 0.0:
            raise ValueError(f"initial_accumulator_value must be non-negative: {initial_accumulator_value}")
        self.initial_accumulator_value = float(initial_accumulator_value)
        self.epsilon = float(epsilon)

    def slot_init(self):
        return (self.initial_accumulator_value,)


class Adam(_Optimizer):
    """b1 = .9, b2 = .999, eps = 1e-7; lr_t = lr*sqrt(1-b2^t)/(1-b1^t)."""
    name = "adam"
    n_slots = 2

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7, **_):
        super().__init__(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)

    def slot_init(self):
        return (0.0, 0.0)

    def lr_t(self, step: int) -> float:
        return self.learning_rate * math.sqrt(1.0 - self.beta_2 ** step) / (1.0 - self.beta_1 ** step)


class OptimizerFactory:
    _supported_optimizers = {"adam": Adam, "adagrad": Adagrad}
    _required_kwargs = ["learning_rate"]

    @classmethod
    def get_optimizer(cls, optimizer_name: str, optimizer_kwargs: Dict[str, Any]) -> _Optimizer:
        if optimizer_name not in cls._supported_optimizers:
            raise ValueError(f"name must be one of {list(cls._supported_optimizers.keys())}, got {optimizer_name}")
        for kwarg in cls._required_kwargs:
            if kwarg not in optimizer_kwargs:
                raise ValueError(f"kwarg {kwarg} not found in kwargs: {optimizer_kwargs}")
        logger.info(f"Creating {optimizer_name} obj with kwargs: {optimizer_kwargs}")
        return cls._supported_optimizers[optimizer_name](**optimizer_kwargs)
