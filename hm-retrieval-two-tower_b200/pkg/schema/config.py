"""Training / architecture settings carried by the Schema (reference pkg/schema/training_config.py:5-39,
pkg/schema/model_config.py:5-25: same field names and defaults)."""
from dataclasses import dataclass
from typing import Any, Dict, List, Optional


@dataclass
class TrainingConfig:
    train_batch_size: int
    test_batch_size: int
    optimizer_name: str                      # resolved by OptimizerFactory
    optimizer_kwargs: Dict[str, Any]         # must hold "learning_rate"
    candidate_batch_size: int = 10000        # rows per candidate-tower pass when building an index
    shuffle_size: Optional[int] = None
    epochs: int = 1
    candidate_prob_lookup: Optional[Dict[str, float]] = None   # id -> sampling probability (logQ)


@dataclass
class ModelConfig:
    joint_embedding_size: int
    ks: List[int]                            # Recall@k cut-offs
    query_tower_units: Optional[List[int]] = None
    candidate_tower_units: Optional[List[int]] = None
