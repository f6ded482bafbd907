"""The one ETL step the hot path depends on: vocabularies and logQ sampling probabilities from the training period
(reference pkg/etl/runner.py:54-84).  The row order of every embedding table (row 0 = OOV, row i + 1 = i-th most frequent id,
features.py:119-127) and the probability table of LogQCorrection (count / len(train), :75-78) are fixed here.

``etl_runner`` -- the join of H&M's raw transactions / articles / customers tables into train.csv and test.csv
(reference :15-51) -- is outside this repository's scope: start from those two CSVs."""
from __future__ import annotations

import logging

from pkg.schema.schema import Schema
from pkg.utils.settings import Settings

logger = logging.getLogger(__name__)


def build_schema_runner(settings: Settings, schema: Schema) -> None:
    import pandas as pd

    logger.info("--- Build Schema Starting ---")
    train = pd.read_csv(settings.train_data_filepath)                    # dtype inference, as load_dataframe (transformations.py:63)
    schema.build_features_from_dataframe(train)
    freq = train[settings.candidate_col_name].value_counts()
    n = len(train)
    # float64 count / n per id, keyed by the id's string form; LogQCorrection casts to fp32 when it builds its table
    lookup = {str(key): count / n for key, count in zip(freq.index, freq.to_numpy(dtype="float64"))}
    logger.info(f"{len(lookup)} candidates with a sampling probability")
    schema.set_candidate_prob_lookup(lookup)
    schema.save(settings.schema_filepath)
    logger.info("--- Build Schema Finished! ---")
