"""Schema: the feature list split by family plus the two config objects (reference pkg/schema/schema.py:13-99)."""
from __future__ import annotations

import logging
import os
import pickle
from typing import Dict, List

from pkg.schema.config import ModelConfig, TrainingConfig
from pkg.schema.features import Feature, FeatureFamily

logger = logging.getLogger(__name__)


class Schema:
    def __init__(self, features: List[Feature], training_config: TrainingConfig, model_config: ModelConfig):
        self.features = features
        self.query_features = [f for f in features if f.feature_family == FeatureFamily.QUERY]
        self.candidate_features = [f for f in features if f.feature_family == FeatureFamily.CANDIDATE]
        self.training_config = training_config
        self.model_config = model_config

    def build_features_from_dataframe(self, df) -> None:
        for f in self.features:
            if not f.is_built:
                f.set_vocab_from_dataframe(df)

    def set_candidate_prob_lookup(self, lookup_dict: Dict[str, float]) -> None:
        logger.info(f"Setting TrainingConfig lookup using dict with {len(lookup_dict)} candidates")
        self.training_config.candidate_prob_lookup = lookup_dict

    def save(self, filepath: str) -> None:
        logger.info(f"Saving Schema obj at filepath: {filepath}")
        os.makedirs(os.path.dirname(filepath) or ".", exist_ok=True)
        with open(filepath, "wb") as fh:
            pickle.dump(self, fh)

    @classmethod
    def load_from_filepath(cls, filepath: str) -> "Schema":
        logger.info(f"Loading Schema obj from {filepath}")
        with open(filepath, "rb") as fh:
            return pickle.load(fh)
