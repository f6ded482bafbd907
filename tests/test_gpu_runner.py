"""End-to-end drop-in check of the TF-free runners (reference pkg/modelling/runner.py:18-152): TFRecords on disk ->
TFRecordDatasetFactory -> TwoTowerModel.fit / BruteForceIndex / IndexRecall per epoch -> saved artefacts."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make_run(tmp_path, epochs=3):
    from pkg.schema import dtypes as tt
    from pkg.schema.config import ModelConfig, TrainingConfig
    from pkg.schema.features import Feature, FeatureFamily
    from pkg.schema.schema import Schema
    from pkg.tfrecord_writer.tfrecord_writer import TFRecordWriter
    from pkg.utils.settings import Settings

    rng = np.random.default_rng(0)
    n_cust, n_art, n_train, n_test = 300, 60, 6000, 1000
    taste = rng.integers(0, n_art, size=n_cust)                      # every customer keeps buying around one article
    def sample(n):
        c = rng.integers(0, n_cust, size=n)
        a = np.where(rng.random(n) < 0.8, taste[c], rng.integers(0, n_art, size=n))
        return {"customer_id": np.array([f"c{i}" for i in c], dtype=object), "article_id": np.array([f"a{i}" for i in a], dtype=object),
                "colour": np.array([f"col{i % 5}" for i in a], dtype=object)}
    train, test = sample(n_train), sample(n_test)
    feats = [Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=16, vocab=sorted(set(train["customer_id"]))),
             Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=16, vocab=sorted(set(train["article_id"]))),
             Feature("colour", tt.string, FeatureFamily.CANDIDATE, embedding_size=4, vocab=[f"col{i}" for i in range(5)])]
    ids, counts = np.unique(train["article_id"], return_counts=True)
    schema = Schema(feats, TrainingConfig(train_batch_size=512, test_batch_size=256, optimizer_name="adagrad", optimizer_kwargs={"learning_rate": 0.1},
                                          candidate_batch_size=32, shuffle_size=2048, epochs=epochs,
                                          candidate_prob_lookup={str(i): float(c) / n_train for i, c in zip(ids, counts)}),
                    ModelConfig(joint_embedding_size=32, ks=[1, 5, 10]))
    d = str(tmp_path)
    s = Settings(raw_data_filepath="", articles_data_filepath="", customers_data_filepath="", train_data_range=("", ""), test_data_range=("", ""),
                 baseline_model_date_range=("", ""), date_col_name="t_dat", candidate_col_name="article_id",
                 candidate_tfrecord_path=f"{d}/cand/candidates.tfrecord", train_data_filepath="", test_data_filepath="",
                 train_data_tfrecord_path=f"{d}/train/train.tfrecord", test_data_tfrecord_path=f"{d}/test/test.tfrecord",
                 schema_filepath=f"{d}/schema.pkl", trained_model_path=f"{d}/model/two_tower", index_path=f"{d}/index/index",
                 baseline_index_path=f"{d}/baseline/index", max_tfrecord_rows=2500)
    schema.save(s.schema_filepath)
    w = TFRecordWriter(schema.features)
    w.write_tfrecords(train, s.train_data_tfrecord_path, s.max_tfrecord_rows)
    w.write_tfrecords(test, s.test_data_tfrecord_path, s.max_tfrecord_rows)
    cand_rows = sorted(set(zip(train["article_id"], train["colour"])) | set(zip(test["article_id"], test["colour"])))     # unique candidates
    TFRecordWriter(schema.candidate_features).write_tfrecords(
        {"article_id": np.array([a for a, _ in cand_rows], dtype=object), "colour": np.array([c for _, c in cand_rows], dtype=object)},
        s.candidate_tfrecord_path, s.max_tfrecord_rows)
    return s, schema, train, test


def test_modelling_runner_end_to_end(tmp_path):
    from pkg.modelling.runner import modelling_runner

    s, schema, train, test = _make_run(tmp_path)
    assert sorted(os.listdir(tmp_path / "train")) == ["train_0.tfrecord", "train_1.tfrecord", "train_2.tfrecord"]
    hist = modelling_runner(s)
    assert len(hist["recall"]) == schema.training_config.epochs + 1 and len(hist["loss"]) == schema.training_config.epochs
    first, last = hist["recall"][0], hist["recall"][-1]
    assert last[10] > first[10] + 0.2 and last[1] > first[1]          # an untrained model is at chance; a trained one finds the taste article
    assert last[1] <= last[5] <= last[10] <= 1.0
    assert hist["loss"][-1] < hist["loss"][0]
    for sub in ("two_tower", "query_tower", "candidate_tower"):        # two_tower_model.py:176-205
        assert os.path.exists(tmp_path / "model" / sub / "variables.npz")
    assert os.path.isdir(tmp_path / "index")


def test_baseline_runner(tmp_path):
    from pkg.modelling.runner import baseline_modelling_runner

    s, schema, train, test = _make_run(tmp_path, epochs=1)
    metric = baseline_modelling_runner(s, candidates=train["article_id"])
    # popularity recall computed by hand from the same data
    ids, counts = np.unique(train["article_id"], return_counts=True)
    order = [i for i, _ in sorted(zip(ids, counts), key=lambda t: -t[1])]
    import pandas as pd

    order = list(pd.Series(train["article_id"]).value_counts().index)   # the reference's tie order
    for k in (1, 5, 10):
        want = np.mean([t in order[:k] for t in test["article_id"]])
        assert abs(metric[k] - want) < 1e-12
