"""Training / evaluation orchestration without TensorFlow (reference pkg/modelling/runner.py:18-152).

Same entry points (``modelling_runner(settings)``, ``baseline_modelling_runner(settings)``) and the same order of work:
per epoch the index is built from the candidate tower AS IT IS BEFORE that epoch's training (reference :88-99), recall
is evaluated over the test set, then the model trains for one epoch and model + index are saved.  TensorBoard callbacks
and summary writers have no counterpart here (metrics go to the log and to the returned history); a final post-training
evaluation -- which the reference lacks: its last ``log_metric`` re-logs the last pre-training numbers (:107) -- is added.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List, Optional

import numpy as np

from pkg.etl.transformations import date_filter, load_dataframe
from pkg.modelling.indices.brute_force import BruteForceIndex
from pkg.modelling.indices.static_index import StaticIndex
from pkg.modelling.losses import CategoricalCrossentropy
from pkg.modelling.metrics.index_recall import IndexRecall
from pkg.modelling.models.two_tower_model import TwoTowerModel
from pkg.modelling.optimizer_factory import OptimizerFactory
from pkg.modelling.tfrecord_dataset import TFRecordDatasetFactory
from pkg.schema.schema import Schema
from pkg.utils.settings import Settings

logger = logging.getLogger(__name__)


def _split_eval(ds, schema: Schema, candidate_col: str):
    """test batches -> (query features, true candidate ids) tuples (reference :49-54)."""
    return ds.map(lambda x: ({f.name: x[f.name] for f in schema.query_features}, x[candidate_col]))


def _candidate_pairs(model: TwoTowerModel, candidate_ds, candidate_col: str):
    """(ids (n,), candidate-tower embeddings (n, E)) per candidate batch (reference :88-93)."""
    for x in candidate_ds:
        yield np.asarray(x[candidate_col]).reshape(-1), model.candidate_tower(x)


def evaluate(model: TwoTowerModel, schema: Schema, candidate_ds, test_ds, candidate_col: str) -> IndexRecall:
    # reference :88-93 maps the candidate tower over the candidate dataset and hands (ids, embeddings) pairs to the index;
    # from_candidate_tower does the same on the device, straight into the corpus buffer (no per-batch host round trip)
    index = BruteForceIndex.from_candidate_tower(max(schema.model_config.ks), model.query_tower, model.candidate_tower, candidate_ds,
                                                 candidate_col)
    metric_calc = IndexRecall(index, schema.model_config.ks)
    for query_features, true_candidates in test_ds:
        metric_calc(query_features, true_candidates)
    return metric_calc


def modelling_runner(settings: Settings, train_ds=None, test_ds=None, candidate_ds=None) -> Dict[str, List]:
    """Train a Two-Tower model and evaluate it each epoch.  The three datasets default to the TFRecord directories named
    in ``settings`` (as the reference does); any re-iterable of {name: (B, 1)} batches may be passed instead."""
    logger.info("--- Modelling Starting ---")
    schema = Schema.load_from_filepath(settings.schema_filepath)
    cfg = schema.training_config
    if train_ds is None:
        train_ds = TFRecordDatasetFactory(schema.features).create_tfrecord_dataset(
            os.path.dirname(settings.train_data_tfrecord_path), cfg.train_batch_size, cfg.shuffle_size)
    if test_ds is None:
        test_ds = TFRecordDatasetFactory(schema.features).create_tfrecord_dataset(
            os.path.dirname(settings.test_data_tfrecord_path), batch_size=cfg.test_batch_size)
    if candidate_ds is None:
        candidate_ds = TFRecordDatasetFactory(schema.candidate_features).create_tfrecord_dataset(
            os.path.dirname(settings.candidate_tfrecord_path), batch_size=cfg.candidate_batch_size)
    test_ds = _split_eval(test_ds, schema, settings.candidate_col_name)
    model = TwoTowerModel.create_from_schema(schema, settings.candidate_col_name)
    optimizer = OptimizerFactory.get_optimizer(cfg.optimizer_name, cfg.optimizer_kwargs)
    model.compile(loss=CategoricalCrossentropy(from_logits=True, reduction="sum"), optimizer=optimizer)
    history: Dict[str, List] = {"recall": [], "loss": []}
    metric_calc: Optional[IndexRecall] = None
    for epoch in range(cfg.epochs):
        metric_calc = evaluate(model, schema, candidate_ds, test_ds, settings.candidate_col_name)   # BEFORE this epoch's training
        history["recall"].append(dict(metric_calc.log_metric(epoch + 1, to_tensorboard=False) or metric_calc.metric))
        h = model.fit(train_ds, epochs=1)
        history["loss"].extend(h["loss"])
        model.save(settings.trained_model_path)
        metric_calc.index.save(settings.index_path)
    # the reference re-logs the last pre-training recall here (:107); evaluate the trained model instead
    metric_calc = evaluate(model, schema, candidate_ds, test_ds, settings.candidate_col_name)
    history["recall"].append(dict(metric_calc.log_metric(cfg.epochs + 1, to_tensorboard=False) or metric_calc.metric))
    metric_calc.index.save(settings.index_path)
    logger.info("--- Modelling Finishing ---")
    return history


def baseline_modelling_runner(settings: Settings, candidates=None, test_ds=None) -> Dict[int, float]:
    """Popularity baseline (reference :111-152).  ``candidates``: the candidate-id column of the training period (a pandas
    Series or any array); when omitted it is read from ``settings.raw_data_filepath`` (CSV) and filtered to
    ``settings.baseline_model_date_range`` like the reference's ETL helpers do."""
    logger.info("--- Baseline Modelling Starting ---")
    schema = Schema.load_from_filepath(settings.schema_filepath)
    if candidates is None:
        # the reference's load_dataframe + date_filter (etl/transformations.py:9-64): dtype inference -- an id column of digits is
        # read as integers, so "0108775015" becomes "108775015" here AND in the TFRecords its ETL writes -- inclusive string dates
        df = load_dataframe(settings.raw_data_filepath, "raw_transactions")
        candidates = date_filter(df, "raw_transactions", settings.date_col_name, settings.baseline_model_date_range)[settings.candidate_col_name]
    if not hasattr(candidates, "value_counts"):
        import pandas as pd

        candidates = pd.Series(np.asarray(candidates).reshape(-1))
    logger.info(f"Building Static Popularity Index using {len(candidates)} candidates")
    if test_ds is None:
        test_ds = TFRecordDatasetFactory(schema.features).create_tfrecord_dataset(
            os.path.dirname(settings.test_data_tfrecord_path), batch_size=schema.training_config.test_batch_size)
    test_ds = _split_eval(test_ds, schema, settings.candidate_col_name)
    index = StaticIndex.build_popularity_index_from_series_schema(schema, candidates)
    metric_calc = IndexRecall(index, schema.model_config.ks)
    for query_features, true_candidates in test_ds:
        metric_calc(query_features, true_candidates)
    metric_calc.log_metric(None, to_tensorboard=False)
    index.save(settings.baseline_index_path)
    logger.info("--- Baseline Modelling Finishing ---")
    return dict(metric_calc.metric)
