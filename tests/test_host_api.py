"""Host-side mirror of the reference API: validation, vocabulary, StaticIndex + IndexRecall (CPU only)."""
import numpy as np
import pytest

from pkg.modelling.indices.static_index import StaticIndex
from pkg.modelling.metrics.index_recall import IndexRecall
from pkg.modelling.optimizer_factory import Adagrad, Adam, OptimizerFactory
from pkg.modelling._device import Vocab
from pkg.schema import dtypes as tt
from pkg.schema.features import Feature, FeatureFamily
from pkg.schema.model_config import ModelConfig
from pkg.schema.schema import Schema
from pkg.schema.training_config import TrainingConfig


def test_feature_validation_matches_reference_errors():
    with pytest.raises(TypeError):
        Feature("x", "int64", FeatureFamily.QUERY)                       # features.py:55-58
    with pytest.raises(ValueError):
        Feature("x", tt.string, "query")                                 # features.py:61-65
    with pytest.raises(TypeError):
        Feature("x", tt.float32, FeatureFamily.QUERY, embedding_size=4)  # features.py:69-72
    with pytest.raises(TypeError):
        Feature("x", tt.string, FeatureFamily.QUERY, max_vocab_size="3")  # features.py:77-80
    f = Feature("age", tt.float32, FeatureFamily.QUERY, vocab=["a"])
    assert f.vocab is None and f.is_built
    g = Feature("id", tt.string, FeatureFamily.CANDIDATE, embedding_size=8)
    assert not g.is_built


def test_vocab_from_dataframe_is_value_counts_order():
    import pandas as pd

    df = pd.DataFrame({"id": ["b", "a", "b", "c", "b", "a"]})
    f = Feature("id", tt.string, FeatureFamily.CANDIDATE, embedding_size=4, max_vocab_size=2)
    f.set_vocab_from_dataframe(df)
    assert list(f.vocab) == ["b", "a"]                                   # features.py:119-127
    with pytest.raises(ValueError):
        Feature("zz", tt.string, FeatureFamily.CANDIDATE, embedding_size=4).set_vocab_from_dataframe(df)


def test_string_lookup_semantics():
    v = Vocab(["query_1", "query_2", "query_3"])
    ids = v.encode(np.array([["query_1"], [b"query_3"], ["query_4"]], dtype=object))
    assert ids.tolist() == [1, 3, 0] and ids.dtype == np.int32 and v.rows == 4


def test_schema_split_and_pickle_roundtrip(tmp_path):
    feats = [Feature("customer_id", tt.string, FeatureFamily.QUERY, 8, vocab=["1"]),
             Feature("article_id", tt.string, FeatureFamily.CANDIDATE, 8, vocab=["2"])]
    s = Schema(feats, TrainingConfig(512, 2048, "adagrad", {"learning_rate": 0.05}), ModelConfig(128, [10, 100]))
    assert [f.name for f in s.query_features] == ["customer_id"]
    assert s.training_config.candidate_batch_size == 10000 and s.training_config.epochs == 1
    s.set_candidate_prob_lookup({"2": 0.5})
    p = tmp_path / "d" / "schema.pkl"
    s.save(str(p))
    s2 = Schema.load_from_filepath(str(p))
    assert s2.training_config.candidate_prob_lookup == {"2": 0.5} and s2.model_config.ks == [10, 100]


def test_optimizer_factory_contract():
    with pytest.raises(ValueError):
        OptimizerFactory.get_optimizer("sgd", {"learning_rate": 0.1})     # optimizer_factory.py:43-48
    with pytest.raises(ValueError):
        OptimizerFactory.get_optimizer("adam", {})                        # optimizer_factory.py:49-53
    a = OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05})
    assert isinstance(a, Adagrad) and a.initial_accumulator_value == 0.1 and a.epsilon == 1e-7
    m = OptimizerFactory.get_optimizer("adam", {"learning_rate": 0.001})
    assert isinstance(m, Adam) and abs(m.lr_t(1) - 0.001 * (1 - 0.999) ** 0.5 / (1 - 0.9)) < 1e-12


def test_recall_with_static_index_reference_fixture(golden):
    fix, _ = golden
    g = fix["recall"]
    feats = [Feature("query_id", tt.string, FeatureFamily.QUERY, embedding_size=2)]
    index = StaticIndex(k=g["static_k"], input_features=feats,
                        candidates=np.array([s.encode() for s in g["static_candidates"]], dtype=object).reshape(1, -1))
    metric = IndexRecall(index, ks=g["ks"])
    q = np.array(g["query_ids"], dtype=object).reshape(-1, 1)
    t = np.array([s.encode() for s in g["true_candidate_ids"]], dtype=object).reshape(-1, 1)
    for i in range(0, 5, g["batch_size"]):
        out = metric({"query_id": q[i:i + 2]}, t[i:i + 2])
    for k, v in g["expected"].items():
        assert metric.metric[int(k)] == np.float64(v) and isinstance(out[int(k)], np.float64)
    assert index({"query_id": q[:3]}).shape == (3, 5)                     # tile(candidates[:, :k], (B, 1))
    assert set(index.get_input_signature()) == {"query_id"}


class _ForeignTensor:
    """Stands in for a TF EagerTensor / CuPy array: exports DLPack and .numpy(), is neither numpy nor torch."""

    def __init__(self, array, dlpack_ok=True):
        self._a, self._ok = array, dlpack_ok
        self.shape = array.shape

    def __dlpack__(self, *args, **kwargs):
        if not self._ok:
            raise TypeError("DT_STRING has no DLPack representation")      # what tf.experimental.dlpack says for strings
        import torch

        return torch.from_numpy(self._a).__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return (1, 0)

    def numpy(self):
        return self._a


def test_foreign_tensors_enter_through_dlpack_or_numpy(golden):
    """north_star: DLPack interop with the TF/Keras tensors the reference feeds its layers (input_layer.py:45-69)."""
    import torch

    from pkg.modelling import _device as D

    ids = D.unwrap(_ForeignTensor(np.arange(6, dtype=np.int32).reshape(6, 1)))
    assert isinstance(ids, torch.Tensor) and ids.dtype == torch.int32 and tuple(ids.shape) == (6, 1)
    strings = _ForeignTensor(np.array([[b"a"], [b"zz"]], dtype=object), dlpack_ok=False)
    assert D.is_string_like(strings) and D.batch_size_of(strings) == 2
    assert list(Vocab(["zz", "a"]).encode(D.unwrap(strings))) == [2, 1]
    for passthrough in (np.zeros(3), torch.zeros(3), [1, 2], 3.0):
        assert D.unwrap(passthrough) is passthrough
    # the reference's recall fixture with the true ids arriving as a string tensor (tests/test_recall.py:48-76)
    fix, _ = golden
    g = fix["recall"]
    index = StaticIndex(k=g["static_k"], input_features=[Feature("query_id", tt.string, FeatureFamily.QUERY, embedding_size=2)],
                        candidates=np.array([s.encode() for s in g["static_candidates"]], dtype=object).reshape(1, -1))
    metric = IndexRecall(index, ks=g["ks"])
    q = np.array(g["query_ids"], dtype=object).reshape(-1, 1)
    t = np.array([s.encode() for s in g["true_candidate_ids"]], dtype=object).reshape(-1, 1)
    for i in range(0, 5, g["batch_size"]):
        metric({"query_id": _ForeignTensor(q[i:i + 2], dlpack_ok=False)}, _ForeignTensor(t[i:i + 2], dlpack_ok=False))
    assert {str(k): float(v) for k, v in metric.metric.items()} == {k: float(v) for k, v in g["expected"].items()}


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pkg._native import TTError
    from pkg.modelling.models.tower import Tower

    with pytest.raises(TTError):
        Tower([Feature("a", tt.string, FeatureFamily.QUERY, 4, vocab=["x"])], 8)


def test_product_never_imports_the_oracle():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hm-retrieval-two-tower_b200")
    bad = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|tt_oracle\.h|libtt_oracle", src, flags=re.M):
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


def test_popularity_index_is_value_counts_order_with_k_max_of_ks():
    """static_index.py:87-95: ids in value_counts() order (descending frequency), k = max(ks), query features as inputs; the
    index then answers every query with the same top-k list (static_index.py:54-55) and feeds IndexRecall unchanged."""
    import pandas as pd

    feats = [Feature("customer_id", tt.string, FeatureFamily.QUERY, 8, vocab=["1"]),
             Feature("article_id", tt.string, FeatureFamily.CANDIDATE, 8, vocab=["2"])]
    schema = Schema(feats, TrainingConfig(512, 2048, "adagrad", {"learning_rate": 0.05}), ModelConfig(128, [1, 3]))
    bought = pd.Series([108, 7, 108, 42, 7, 108, 9])                     # integer ids become strings, as the reference does
    index = StaticIndex.build_popularity_index_from_series_schema(schema, bought)
    assert index.k == 3 and [f.name for f in index.input_features] == ["customer_id"]
    assert index.candidates.shape == (1, 4) and list(index.candidates[0, :2]) == ["108", "7"]
    out = index({"customer_id": np.array([["a"], ["b"]], dtype=object)})
    assert out.shape == (2, 3) and (out[0] == out[1]).all() and list(out[0][:2]) == ["108", "7"]
    metric = IndexRecall(index, ks=[1, 3])
    got = metric({"customer_id": np.array([["a"], ["b"], ["c"]], dtype=object)}, np.array([["108"], ["7"], ["nope"]], dtype=object))
    assert got[1] == np.float64(1) / 3 and got[3] == np.float64(2) / 3 and metric.seen == 3


def test_logq_host_lookup_on_the_reference_fixture(golden):
    """Host half of LogQCorrection on the reference's fixture (tests/test_layers.py:8-36): string -> probability with default 1.0
    (logq_correction.py:32-42), per-row ln p for the fused path; ln and the subtraction are the GPU half (tests/test_gpu_model.py)."""
    from pkg.modelling.layers.logq_correction import LogQCorrection

    fix, _ = golden
    g = fix["logq"]
    layer = LogQCorrection(g["candidate_prob_lookup"])
    ids = np.array([[s.encode()] for s in g["candidate_ids"]], dtype=object)
    p = layer.probabilities(ids)
    assert p.dtype == np.float32 and p.tolist() == [np.float32(0.3), np.float32(0.2), np.float32(0.5)]
    np.testing.assert_allclose(np.asarray(g["logits"], np.float32) - np.log(p)[None, :], g["expected"], rtol=1e-6)
    assert layer.probabilities(np.array([["id9"], ["id2"]], dtype=object)).tolist() == [1.0, np.float32(0.2)]   # unknown id: p = 1
    rows = layer.row_probabilities(Vocab(["id3", "id1", "zz"]))        # row 0 = OOV, then vocabulary order; id2 is outside
    assert rows.tolist() == [1.0, np.float32(0.5), np.float32(0.3), 1.0]


def test_abstract_model_signature_defaults_and_save(tmp_path):
    """abstract_keras_model.py:10-131 without TensorFlow: (None, 1) signature per feature, default (1, 1) inputs per dtype,
    TypeError for any other dtype (:63-68), save() writes the model's arrays."""
    from pkg.modelling.models.abstract_keras_model import AbstractKerasModel, TensorSpec

    feats = [Feature("query_id", tt.string, FeatureFamily.QUERY, embedding_size=2), Feature("age", tt.float32, FeatureFamily.QUERY)]
    index = StaticIndex(k=2, input_features=feats, candidates=np.array([["a", "b", "c"]], dtype=object))
    sig = index.get_input_signature()
    assert sig == {"query_id": TensorSpec((None, 1), tt.string, "query_id"), "age": TensorSpec((None, 1), tt.float32, "age")}
    assert index._input_signature == sig                                  # recorded by initialise_model() in the constructor
    d = index.get_default_inputs(sig)
    assert d["query_id"].shape == (1, 1) and d["query_id"][0, 0] == "a" and d["age"].dtype == np.float32 and d["age"][0, 0] == 0.0
    with pytest.raises(TypeError):
        AbstractKerasModel._get_default_tensor("int64")
    index.save(str(tmp_path / "m" / "static"))
    with np.load(tmp_path / "m" / "static" / "variables.npz") as z:
        assert z["candidates"].tolist() == [["a", "b", "c"]]
    with pytest.raises(TypeError):
        AbstractKerasModel()                                              # abstract: call / get_input_signature must be provided
