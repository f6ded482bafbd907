"""Entry point in the shape of the reference's main.py (main.py:13-127): Settings + Schema, then the runners in order.

    python main.py --data-dir ./data                          # H&M's transactions_train.csv / articles.csv / customers.csv in ./data
    python main.py --data-dir /tmp/run --synthetic 200000     # H&M-shaped synthetic raw tables instead
    python main.py --data-dir ./data --steps schema,tfrecords,baseline      # any subset; `model` needs a B200

Steps, in the reference's order (main.py:123-127): ``etl`` (raw tables -> train.csv / test.csv, etl/runner.py:15-51), ``schema``
(vocabularies + logQ table, etl/runner.py:54-84), ``tfrecords`` (tfrecord_writer/runner.py:13-60), ``model``
(modelling/runner.py:18-108) and ``baseline`` (modelling/runner.py:111-152)."""
from __future__ import annotations

import argparse
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from pkg.schema import dtypes as tt  # noqa: E402
from pkg.schema.config import ModelConfig, TrainingConfig  # noqa: E402
from pkg.schema.features import Feature, FeatureFamily  # noqa: E402
from pkg.schema.schema import Schema  # noqa: E402
from pkg.utils.settings import Settings  # noqa: E402

logger = logging.getLogger("pkg.main")
STEPS = ("etl", "schema", "tfrecords", "model", "baseline")


def make_settings(data_dir: str, max_tfrecord_rows: int = 100000) -> Settings:
    d = data_dir
    return Settings(
        raw_data_filepath=f"{d}/transactions_train.csv", articles_data_filepath=f"{d}/articles.csv", customers_data_filepath=f"{d}/customers.csv",
        train_data_range=("2019-09-20", "2020-08-20"), test_data_range=("2020-08-21", "2020-09-21"),
        baseline_model_date_range=("2019-09-20", "2020-08-20"), date_col_name="t_dat", candidate_col_name="article_id",
        candidate_tfrecord_path=f"{d}/tfrecords/candidates/candidates", train_data_filepath=f"{d}/train.csv", test_data_filepath=f"{d}/test.csv",
        train_data_tfrecord_path=f"{d}/tfrecords/train/train", test_data_tfrecord_path=f"{d}/tfrecords/test/test",
        max_tfrecord_rows=max_tfrecord_rows, schema_filepath=f"{d}/schema.pkl", trained_model_path=f"{d}/trained_models/model/",
        index_path=f"{d}/trained_models/candidate_index", baseline_index_path=f"{d}/trained_models/baseline_index")


def make_schema(joint: int = 64, batch: int = 8192, epochs: int = 1) -> Schema:
    """BASELINE.json configs[1]: id towers of width ``joint`` with article side features; Adagrad lr 0.05 as main.py:100-101."""
    return Schema(
        features=[
            Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=joint),
            Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=joint),
            Feature("product_type_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=16),
            Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=8),
        ],
        training_config=TrainingConfig(train_batch_size=batch, test_batch_size=2048, optimizer_name="adagrad",
                                       optimizer_kwargs={"learning_rate": 0.05}, shuffle_size=100000, epochs=epochs),
        model_config=ModelConfig(joint_embedding_size=joint, ks=[10, 100]))


def write_synthetic(settings: Settings, rows: int, customers: int = 20000, articles: int = 2000, seed: int = 0) -> None:
    """H&M-shaped raw tables: transactions (uniform customers, Zipf-popular articles, a per-customer taste so that a trained model
    can beat popularity), articles with side features that are functions of the article, customers."""
    import numpy as np
    import pandas as pd

    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, articles + 1)
    taste = rng.choice(articles, size=customers, p=pop / pop.sum())
    c = rng.integers(0, customers, rows)
    a = np.where(rng.random(rows) < 0.6, taste[c], rng.choice(articles, size=rows, p=pop / pop.sum()))
    lo, hi = pd.Timestamp(settings.train_data_range[0]), pd.Timestamp(settings.test_data_range[1])
    day = lo + pd.to_timedelta(np.sort(rng.integers(0, (hi - lo).days + 1, rows)), unit="D")
    os.makedirs(os.path.dirname(settings.raw_data_filepath) or ".", exist_ok=True)
    pd.DataFrame({"t_dat": day.strftime("%Y-%m-%d"), "customer_id": [f"cust{i:07d}" for i in c], "article_id": 100000000 + a}).to_csv(
        settings.raw_data_filepath, index=False)
    ids = np.arange(articles)
    pd.DataFrame({"article_id": 100000000 + ids, "product_type_name": [f"type{i % 131}" for i in ids],
                  "colour_group_name": [f"colour{i % 50}" for i in ids]}).to_csv(settings.articles_data_filepath, index=False)
    pd.DataFrame({"customer_id": [f"cust{i:07d}" for i in range(customers)], "age": rng.integers(16, 90, customers)}).to_csv(
        settings.customers_data_filepath, index=False)


def run(settings: Settings, schema: Schema, steps) -> dict:
    out = {}
    if "etl" in steps:
        from pkg.etl.runner import etl_runner

        etl_runner(settings)
    if "schema" in steps:
        from pkg.etl.runner import build_schema_runner

        build_schema_runner(settings, schema)
    if "tfrecords" in steps:
        from pkg.tfrecord_writer.runner import tfrecord_writer_runner

        tfrecord_writer_runner(settings)
    if "model" in steps:
        from pkg.modelling.runner import modelling_runner

        out["model"] = modelling_runner(settings)
    if "baseline" in steps:
        from pkg.modelling.runner import baseline_modelling_runner

        out["baseline"] = baseline_modelling_runner(settings)
    return out


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data-dir", default="./data")
    ap.add_argument("--synthetic", type=int, default=0, metavar="ROWS", help="write synthetic raw tables with ROWS transactions into --data-dir first")
    ap.add_argument("--steps", default=",".join(STEPS), help="comma-separated subset of " + ",".join(STEPS))
    ap.add_argument("--joint", type=int, default=64)
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--max-tfrecord-rows", type=int, default=100000)
    args = ap.parse_args(argv)
    steps = [s for s in args.steps.split(",") if s]
    unknown = [s for s in steps if s not in STEPS]
    if unknown:
        raise ValueError(f"unknown step(s) {unknown}; choose from {list(STEPS)}")
    settings = make_settings(args.data_dir, args.max_tfrecord_rows)
    if args.synthetic:
        write_synthetic(settings, args.synthetic)
    result = run(settings, make_schema(args.joint, args.batch, args.epochs), steps)
    for name, value in result.items():
        logger.info(f"{name}: {value}")
    return result


if __name__ == "__main__":
    main()
