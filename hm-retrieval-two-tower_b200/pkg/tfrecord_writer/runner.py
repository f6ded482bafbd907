"""CSV -> TFRecords step between the ETL and the modelling runner (reference pkg/tfrecord_writer/runner.py:13-60), TF-free.

``tfrecord_writer_runner(settings)`` leaves three sets of partitioned files behind: the unique candidate rows of train + test
(the corpus the index is built from and evaluated against), the train rows and the test rows.  CSVs are read with pandas' dtype
inference exactly as the reference's ``load_dataframe`` does (etl/transformations.py:63), so an all-digit id column is written
without its leading zeros -- consistently on every path that later looks those ids up."""
from __future__ import annotations

import logging

from pkg.etl.transformations import load_dataframe
from pkg.schema.schema import Schema
from pkg.tfrecord_writer.tfrecord_writer import TFRecordWriter
from pkg.utils.settings import Settings

logger = logging.getLogger(__name__)


def tfrecord_writer_runner(settings: Settings) -> None:
    import pandas as pd

    logger.info("--- TFRecord Writing Starting ---")
    schema = Schema.load_from_filepath(settings.schema_filepath)
    frames = {"train": load_dataframe(settings.train_data_filepath, "train"), "test": load_dataframe(settings.test_data_filepath, "test")}
    # every candidate seen in either period, one row each (a candidate is assumed to carry the same side features everywhere)
    cols = [f.name for f in schema.candidate_features]
    seen = pd.concat(list(frames.values()))[cols]
    unique = seen.drop_duplicates()
    logger.info(f"{len(seen)} candidate rows in train + test, {len(unique)} unique")
    TFRecordWriter(schema.candidate_features).write_tfrecords(unique, settings.candidate_tfrecord_path, settings.max_tfrecord_rows)
    writer = TFRecordWriter(schema.features)
    for name, path in (("train", settings.train_data_tfrecord_path), ("test", settings.test_data_tfrecord_path)):
        writer.write_tfrecords(frames[name], path, settings.max_tfrecord_rows)
    logger.info("--- TFRecord Writing Finishing ---")
