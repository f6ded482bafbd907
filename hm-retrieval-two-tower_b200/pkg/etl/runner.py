"""ETL steps in front of the hot path (reference pkg/etl/runner.py:15-84).

``build_schema_runner`` is the one the path depends on: it fixes the row order of every embedding table (row 0 = OOV, row i + 1 =
i-th most frequent id, features.py:119-127) and the probability table of LogQCorrection (count / len(train), :75-78).
``etl_runner`` joins the raw transactions / articles / customers tables and cuts train.csv / test.csv by date (:15-51)."""
from __future__ import annotations

import logging

from pkg.etl.transformations import date_filter, load_dataframe, save_dataframe
from pkg.schema.schema import Schema
from pkg.utils.settings import Settings

logger = logging.getLogger(__name__)


def etl_runner(settings: Settings) -> None:
    logger.info("--- ETL Starting ---")
    joined = load_dataframe(settings.raw_data_filepath, "raw_transactions")
    for path, name, key in ((settings.articles_data_filepath, "articles", "article_id"), (settings.customers_data_filepath, "customers", "customer_id")):
        joined = joined.merge(load_dataframe(path, name), how="inner", on=key)      # transactions without metadata are dropped
    for name, span, path in (("train", settings.train_data_range, settings.train_data_filepath),
                             ("test", settings.test_data_range, settings.test_data_filepath)):
        save_dataframe(date_filter(joined, name, settings.date_col_name, span), name, settings.date_col_name, path)
    logger.info("--- ETL Finished! ---")


def build_schema_runner(settings: Settings, schema: Schema) -> None:
    logger.info("--- Build Schema Starting ---")
    train = load_dataframe(settings.train_data_filepath, "train")
    schema.build_features_from_dataframe(train)
    freq = train[settings.candidate_col_name].value_counts()
    n = len(train)
    # float64 count / n per id, keyed by the id's string form; LogQCorrection casts to fp32 when it builds its table
    lookup = {str(key): count / n for key, count in zip(freq.index, freq.to_numpy(dtype="float64"))}
    logger.info(f"{len(lookup)} candidates with a sampling probability")
    schema.set_candidate_prob_lookup(lookup)
    schema.save(settings.schema_filepath)
    logger.info("--- Build Schema Finished! ---")
