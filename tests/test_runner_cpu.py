"""baseline_modelling_runner end to end on the CPU (reference pkg/modelling/runner.py:111-152): raw transactions CSV -> date filter ->
popularity StaticIndex -> Recall@k over the test TFRecords -> saved index.  Nothing on this path is numeric, so it needs no GPU."""
import os

import numpy as np

from pkg.modelling.runner import baseline_modelling_runner
from pkg.schema import dtypes as tt
from pkg.schema.config import ModelConfig, TrainingConfig
from pkg.schema.features import Feature, FeatureFamily
from pkg.schema.schema import Schema
from pkg.tfrecord_writer.tfrecord_writer import TFRecordWriter
from pkg.utils.settings import Settings


def test_tfrecord_writer_runner_then_baseline(tmp_path):
    """tfrecord_writer/runner.py:13-60 followed by the baseline runner, as main.py chains them: train / test CSVs -> candidate,
    train and test TFRecord partitions -> popularity recall.  Ids with leading zeros survive consistently (both sides drop them)."""
    import pandas as pd

    from pkg.modelling.tfrecord_dataset import TFRecordDatasetFactory
    from pkg.tfrecord_writer.runner import tfrecord_writer_runner

    rng = np.random.default_rng(9)
    arts = np.array(["0108775015", "0108775044", "0110065001", "0111565001", "0111586001"])
    colour = {a: f"col{i % 2}" for i, a in enumerate(arts)}

    def period(n, lo, hi):
        a = rng.choice(arts, n, p=[0.4, 0.25, 0.15, 0.12, 0.08])
        return pd.DataFrame({"t_dat": rng.choice(pd.date_range(lo, hi).strftime("%Y-%m-%d"), n),
                             "customer_id": [f"c{i}" for i in rng.integers(0, 30, n)], "article_id": a, "colour": [colour[x] for x in a]})

    d = str(tmp_path)
    train, test = period(230, "2020-09-01", "2020-09-15"), period(70, "2020-09-16", "2020-09-22")
    train.to_csv(f"{d}/train.csv", index=False)
    test.to_csv(f"{d}/test.csv", index=False)
    pd.concat([train, test]).to_csv(f"{d}/transactions.csv", index=False)
    feats = [Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=4, vocab=["c0"]),
             Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=4, vocab=["108775015"]),
             Feature("colour", tt.string, FeatureFamily.CANDIDATE, embedding_size=2, vocab=["col0", "col1"])]
    schema = Schema(feats, TrainingConfig(train_batch_size=64, test_batch_size=32, optimizer_name="adagrad", optimizer_kwargs={"learning_rate": 0.1}),
                    ModelConfig(joint_embedding_size=8, ks=[1, 2]))
    s = Settings(raw_data_filepath=f"{d}/transactions.csv", articles_data_filepath="", customers_data_filepath="",
                 train_data_range=("2020-09-01", "2020-09-15"), test_data_range=("2020-09-16", "2020-09-22"),
                 baseline_model_date_range=("2020-09-01", "2020-09-15"), date_col_name="t_dat", candidate_col_name="article_id",
                 candidate_tfrecord_path=f"{d}/cand/candidates.tfrecord", train_data_filepath=f"{d}/train.csv", test_data_filepath=f"{d}/test.csv",
                 train_data_tfrecord_path=f"{d}/train/train.tfrecord", test_data_tfrecord_path=f"{d}/test/test.tfrecord",
                 schema_filepath=f"{d}/schema.pkl", trained_model_path=f"{d}/model/m", index_path=f"{d}/index/i",
                 baseline_index_path=f"{d}/baseline/index", max_tfrecord_rows=100)
    schema.save(s.schema_filepath)
    tfrecord_writer_runner(s)
    assert sorted(os.listdir(f"{d}/train")) == ["train_0.tfrecord", "train_1.tfrecord", "train_2.tfrecord"]       # 100 + 100 + 30 rows
    assert os.listdir(f"{d}/test") == ["test_0.tfrecord"] and os.listdir(f"{d}/cand") == ["candidates_0.tfrecord"]
    cand = list(TFRecordDatasetFactory(schema.candidate_features).create_tfrecord_dataset(f"{d}/cand", batch_size=100))[0]
    got = sorted((a.decode(), c.decode()) for a, c in zip(cand["article_id"].reshape(-1), cand["colour"].reshape(-1)))
    assert got == sorted((str(int(a)), colour[a]) for a in arts)                 # five unique candidates, ids as pandas inferred them
    rows = sum(b["customer_id"].shape[0] for b in TFRecordDatasetFactory(schema.features).create_tfrecord_dataset(f"{d}/train", batch_size=64))
    assert rows == 230
    metric = baseline_modelling_runner(s)
    order = [str(a) for a in pd.read_csv(f"{d}/train.csv").article_id.value_counts().index]
    truth = [str(int(a)) for a in test.article_id]
    for k in (1, 2):
        assert metric[k] == np.float64(sum(t in order[:k] for t in truth)) / np.float64(len(truth)) and metric[k] > 0


def test_baseline_runner_from_csv_and_tfrecords(tmp_path):
    import pandas as pd

    rng = np.random.default_rng(5)
    n = 400
    days = pd.date_range("2020-08-01", "2020-09-22").strftime("%Y-%m-%d")
    popular = np.array([108775015, 108775044, 110065001, 111565001, 111586001, 111593001, 111609001])
    weights = np.array([30, 20, 15, 12, 10, 8, 5], dtype=np.float64)
    raw = pd.DataFrame({"t_dat": rng.choice(days, n), "customer_id": [f"c{i}" for i in rng.integers(0, 50, n)],
                        "article_id": [f"0{a}" for a in rng.choice(popular, n, p=weights / weights.sum())]})   # leading zero, as in H&M's file
    d = str(tmp_path)
    raw.to_csv(f"{d}/transactions.csv", index=False)
    feats = [Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=4, vocab=["c0"]),
             Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=4, vocab=["108775015"])]
    schema = Schema(feats, TrainingConfig(train_batch_size=64, test_batch_size=32, optimizer_name="adagrad", optimizer_kwargs={"learning_rate": 0.1}),
                    ModelConfig(joint_embedding_size=8, ks=[1, 3, 5]))
    s = Settings(raw_data_filepath=f"{d}/transactions.csv", articles_data_filepath="", customers_data_filepath="",
                 train_data_range=("2020-08-01", "2020-09-15"), test_data_range=("2020-09-16", "2020-09-22"),
                 baseline_model_date_range=("2020-09-01", "2020-09-15"), date_col_name="t_dat", candidate_col_name="article_id",
                 candidate_tfrecord_path=f"{d}/cand/c.tfrecord", train_data_filepath="", test_data_filepath="",
                 train_data_tfrecord_path=f"{d}/train/train.tfrecord", test_data_tfrecord_path=f"{d}/test/test.tfrecord",
                 schema_filepath=f"{d}/schema.pkl", trained_model_path=f"{d}/model/m", index_path=f"{d}/index/i",
                 baseline_index_path=f"{d}/baseline/index")
    schema.save(s.schema_filepath)
    # the test period as the reference's ETL would write it: the CSV re-read with pandas' dtype inference, ids stringified
    df = pd.read_csv(s.raw_data_filepath)
    test = df[(df.t_dat >= s.test_data_range[0]) & (df.t_dat <= s.test_data_range[1])]
    assert len(test) > 20 and df.article_id.dtype.kind == "i"
    TFRecordWriter(schema.features).write_tfrecords(
        {"customer_id": test.customer_id.to_numpy(dtype=object), "article_id": np.array([str(a) for a in test.article_id], dtype=object)},
        s.test_data_tfrecord_path)
    metric = baseline_modelling_runner(s)
    window = df[(df.t_dat >= "2020-09-01") & (df.t_dat <= "2020-09-15")].article_id       # both ends inclusive
    order = [str(a) for a in window.value_counts().index]
    assert not order[0].startswith("0")
    for k in (1, 3, 5):
        want = np.float64(sum(str(a) in order[:k] for a in test.article_id)) / np.float64(len(test))
        assert metric[k] == want and 0 < want <= 1
    assert metric[1] <= metric[3] <= metric[5]
    with np.load(os.path.join(s.baseline_index_path, "variables.npz")) as z:
        assert z["candidates"].reshape(-1).tolist() == order


def test_build_schema_runner_vocabulary_order_and_logq_table(tmp_path):
    """etl/runner.py:54-84: vocabularies in value_counts order (row i + 1 = i-th most frequent id, features.py:119-127) and
    p(id) = count / len(train) keyed by str(id) (:75-78)."""
    import pandas as pd

    from pkg.etl.runner import build_schema_runner

    d = str(tmp_path)
    train = pd.DataFrame({"customer_id": ["b", "a", "b", "c", "b", "a"], "article_id": [7, 7, 9, 7, 108, 9]})
    train.to_csv(f"{d}/train.csv", index=False)
    feats = [Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=4),
             Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=4, max_vocab_size=2)]
    schema = Schema(feats, TrainingConfig(8, 8, "adagrad", {"learning_rate": 0.05}), ModelConfig(8, [1]))
    s = Settings(raw_data_filepath="", articles_data_filepath="", customers_data_filepath="", train_data_range=("", ""), test_data_range=("", ""),
                 baseline_model_date_range=("", ""), date_col_name="t_dat", candidate_col_name="article_id", candidate_tfrecord_path="",
                 train_data_filepath=f"{d}/train.csv", test_data_filepath="", train_data_tfrecord_path="", test_data_tfrecord_path="",
                 schema_filepath=f"{d}/out/schema.pkl", trained_model_path="", index_path="", baseline_index_path="")
    build_schema_runner(s, schema)
    loaded = Schema.load_from_filepath(s.schema_filepath)
    assert list(loaded.features[0].vocab) == ["b", "a", "c"] and list(loaded.features[1].vocab) == ["7", "9"]     # max_vocab_size = 2
    assert loaded.training_config.candidate_prob_lookup == {"7": 3 / 6, "9": 2 / 6, "108": 1 / 6}
    assert all(f.is_built for f in loaded.features)


def test_main_entry_point_cpu_steps(tmp_path):
    """main.py-shaped entry (reference main.py:13-127) on synthetic raw tables: ETL -> schema -> TFRecords -> popularity baseline."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tt_main", os.path.join(root, "hm-retrieval-two-tower_b200", "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(["--data-dir", str(tmp_path), "--synthetic", "3000", "--steps", "etl,schema,tfrecords,baseline", "--max-tfrecord-rows", "1000"])
    recall = out["baseline"]
    assert set(recall) == {10, 100} and 0 < recall[10] <= recall[100] <= 1
    schema = Schema.load_from_filepath(str(tmp_path / "schema.pkl"))
    assert [f.name for f in schema.candidate_features] == ["article_id", "product_type_name", "colour_group_name"]
    probs = schema.training_config.candidate_prob_lookup
    assert abs(sum(probs.values()) - 1.0) < 1e-9 and all(k.isdigit() for k in probs)
    assert len(os.listdir(tmp_path / "tfrecords" / "train")) == 3 and os.path.exists(tmp_path / "trained_models" / "baseline_index" / "variables.npz")
    import pytest

    with pytest.raises(ValueError):
        mod.main(["--data-dir", str(tmp_path), "--steps", "bogus"])


def test_date_filter_like_the_reference_test():
    """The reference's tests/test_transformations.py:7-38 on the same frame: both ends of the range are inclusive."""
    import pandas as pd

    from pkg.etl.transformations import date_filter

    df = pd.DataFrame({"date_col": ["2024-01-01", "2024-02-01", "2024-03-01", "2024-04-01", "2024-05-01"],
                       "query_id": ["123", "456", "123", "789", "456"], "candidate_id": ["abc", "def", "ghi", "abc", "def"]})
    train = date_filter(df, "train", "date_col", ("2024-01-01", "2024-02-01"))
    test = date_filter(df, "train", "date_col", ("2024-03-01", "2024-04-01"))
    assert (train["date_col"].min(), train["date_col"].max()) == ("2024-01-01", "2024-02-01")
    assert (test["date_col"].min(), test["date_col"].max()) == ("2024-03-01", "2024-04-01")
    assert len(date_filter(df, "none", "date_col", ("2025-01-01", "2025-02-01"))) == 0


def test_etl_runner_joins_and_splits(tmp_path):
    """etl/runner.py:15-51: inner joins with the article and customer tables, then the two date windows."""
    import pandas as pd

    from pkg.etl.runner import etl_runner

    d = str(tmp_path)
    pd.DataFrame({"t_dat": ["2020-01-01", "2020-01-05", "2020-02-01", "2020-02-03", "2020-03-01"], "customer_id": ["a", "b", "a", "zz", "b"],
                  "article_id": ["0101", "0102", "0101", "0102", "0999"]}).to_csv(f"{d}/tx.csv", index=False)
    pd.DataFrame({"article_id": ["0101", "0102"], "colour": ["red", "blue"]}).to_csv(f"{d}/articles.csv", index=False)
    pd.DataFrame({"customer_id": ["a", "b"], "age": [30, 40]}).to_csv(f"{d}/customers.csv", index=False)
    s = Settings(raw_data_filepath=f"{d}/tx.csv", articles_data_filepath=f"{d}/articles.csv", customers_data_filepath=f"{d}/customers.csv",
                 train_data_range=("2020-01-01", "2020-01-31"), test_data_range=("2020-02-01", "2020-03-31"), baseline_model_date_range=("", ""),
                 date_col_name="t_dat", candidate_col_name="article_id", candidate_tfrecord_path="", train_data_filepath=f"{d}/out/train.csv",
                 test_data_filepath=f"{d}/out/test.csv", train_data_tfrecord_path="", test_data_tfrecord_path="", schema_filepath="",
                 trained_model_path="", index_path="", baseline_index_path="")
    etl_runner(s)
    train, test = pd.read_csv(s.train_data_filepath), pd.read_csv(s.test_data_filepath)
    assert train[["customer_id", "article_id", "colour", "age"]].values.tolist() == [["a", 101, "red", 30], ["b", 102, "blue", 40]]
    assert test[["customer_id", "article_id"]].values.tolist() == [["a", 101]]      # unknown customer zz and unknown article 0999 are dropped
