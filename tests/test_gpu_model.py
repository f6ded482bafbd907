"""The reference-facing Python API (pkg.modelling) on the GPU, checked against the oracle, the reference's
own fixtures and the KAT vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import two_tower_oracle as O  # noqa: E402
from pkg.schema import dtypes as tt  # noqa: E402
from pkg.schema.features import Feature, FeatureFamily  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _set(t, a):
    import torch

    t.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda())


def _kat_model(a, lib):
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory
    from pkg.modelling.losses import CategoricalCrossentropy, Reduction

    qf = [Feature("q", tt.string, FeatureFamily.QUERY, embedding_size=2, vocab=["1", "2", "3"])]
    cf = [Feature("c", tt.string, FeatureFamily.CANDIDATE, embedding_size=2, vocab=["1", "2"])]
    p = a["p_row"]
    m = TwoTowerModel(qf, cf, "c", 2, candidate_prob_lookup={"1": p[1], "2": p[2]})
    _set(m.query_tower.input_layer.embedding_layers["q"].weight, a["Tq"])
    _set(m.candidate_tower.input_layer.embedding_layers["c"].weight, a["Tc"])
    _set(m.query_tower.kernels[0], a["Wq"]); _set(m.query_tower.biases[0], a["bq"])
    _set(m.candidate_tower.kernels[0], a["Wc"]); _set(m.candidate_tower.biases[0], a["bc"])
    m.compile(loss=CategoricalCrossentropy(from_logits=True, reduction=Reduction.SUM),
              optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": a["lr"]}))
    return m


@pytest.mark.parametrize("encoded", [False, True])
def test_kat_a_train_step_through_the_api(lib, golden, encoded):
    _, kat = golden
    a = kat["A"]
    m = _kat_model(a, lib)
    if encoded:   # pre-encoded row ids (fast path) -- same result as strings
        data = {"q": np.array(a["query_ids"], np.int32).reshape(-1, 1), "c": np.array(a["candidate_ids"], np.int32).reshape(-1, 1)}
    else:
        data = {"q": np.array([[str(i)] for i in a["query_ids"]], dtype=object),
                "c": np.array([[str(i)] for i in a["candidate_ids"]], dtype=object)}
    logits = _np(m(data))
    np.testing.assert_allclose(logits, a["S"], atol=1e-7)                       # call(): raw Q.C^T, no logQ (:65-92)
    out = m.train_step(data)
    assert abs(float(out["loss"]) - a["loss"]) < 1e-5
    np.testing.assert_allclose(_np(m.candidate_tower.input_layer.embedding_layers["c"].weight), a["Tc_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m.query_tower.input_layer.embedding_layers["q"].weight), a["Tq_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m.query_tower.kernels[0]), a["Wq_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m.query_tower.biases[0]), a["bq_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m.candidate_tower.kernels[0]), a["Wc_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m.candidate_tower.biases[0]), a["bc_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m._opt_state["tables"][id(m.candidate_tower.input_layer.embedding_layers["c"])][0]),
                               a["acc_c_after"], atol=1e-6)


class _DLPackOnly:
    """A device tensor as a TF EagerTensor / CuPy array would present it: DLPack export and a shape, nothing else."""

    def __init__(self, t):
        self._t, self.shape = t, tuple(t.shape)

    def __dlpack__(self, *args, **kwargs):
        return self._t.__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_kat_a_train_step_with_dlpack_inputs(lib, golden):
    """Row ids handed over as foreign device tensors (DLPack) give the KAT-A step of the string path."""
    import torch

    _, kat = golden
    a = kat["A"]
    m = _kat_model(a, lib)
    data = {"q": _DLPackOnly(torch.tensor(a["query_ids"], dtype=torch.int32, device="cuda").reshape(-1, 1)),
            "c": _DLPackOnly(torch.tensor(a["candidate_ids"], dtype=torch.int32, device="cuda").reshape(-1, 1))}
    np.testing.assert_allclose(_np(m(data)), a["S"], atol=1e-7)
    out = m.train_step(data)
    assert abs(float(out["loss"]) - a["loss"]) < 1e-5
    np.testing.assert_allclose(_np(m.candidate_tower.input_layer.embedding_layers["c"].weight), a["Tc_after"], atol=1e-6)
    np.testing.assert_allclose(_np(m.query_tower.kernels[0]), a["Wq_after"], atol=1e-6)


def test_constructor_errors(lib):
    from pkg.modelling.models.two_tower_model import TwoTowerModel

    qf = [Feature("q", tt.string, FeatureFamily.QUERY, embedding_size=2, vocab=["1"])]
    cf = [Feature("c", tt.string, FeatureFamily.CANDIDATE, embedding_size=2, vocab=["1"])]
    with pytest.raises(ValueError):
        TwoTowerModel(qf, cf, "nope", 2)                                         # two_tower_model.py:47-50
    m = TwoTowerModel(qf, cf, "c", 2)
    with pytest.raises(RuntimeError):
        m.train_step({"q": np.array([["1"]]), "c": np.array([["1"]])})
    assert set(m.get_input_signature()) == {"q", "c"}
    assert m.get_input_signature()["q"].shape == (None, 1)


def test_logq_layer_reference_fixture(lib, golden):
    from pkg.modelling.layers.logq_correction import LogQCorrection

    fix, _ = golden
    g = fix["logq"]
    layer = LogQCorrection(g["candidate_prob_lookup"])
    out = layer(np.array(g["logits"], np.float32), np.array(g["candidate_ids"]).reshape(3, 1))
    np.testing.assert_allclose(_np(out), np.array(g["expected"]), rtol=0, atol=5e-7)   # tests/test_layers.py:26-39
    same = layer(np.array(g["logits"], np.float32), np.array(["zz", "zz", "zz"]).reshape(3, 1))
    np.testing.assert_allclose(_np(same), np.array(g["logits"]), atol=0)        # unknown ids: ln(1) = 0


class _MockEmbeddingModel:
    """The reference test's fake query tower (tests/test_indices.py:8-60): StringLookup + row gather."""

    def __init__(self, vocab, table):
        from pkg.modelling._device import Vocab

        self.vocab, self.table = Vocab(vocab), np.asarray(table, np.float32)

    def __call__(self, x):
        return self.table[self.vocab.encode(x["id"])]

    def get_input_signature(self):
        from pkg.modelling.models.abstract_keras_model import TensorSpec

        return {"id": TensorSpec((None, 1), tt.string, "id")}


def test_brute_force_index_reference_fixture(lib, golden):
    from pkg.modelling.indices.brute_force import BruteForceIndex
    from pkg.modelling.metrics.index_recall import IndexRecall

    fix, _ = golden
    g = fix["brute_force"]
    qm = _MockEmbeddingModel(g["query_vocab"], g["query_table"])
    pairs = [(np.array([i]), np.array([e], np.float32)) for i, e in zip(g["candidate_ids"], g["candidate_embeddings"])]  # batch(1)
    index = BruteForceIndex(g["k"], qm, pairs)
    inputs = {"id": np.array(g["queries"], dtype=object).reshape(5, 1)}
    assert index(inputs).tolist() == g["expected"]                               # tests/test_indices.py:105-132
    assert index._candidates.shape == (5, 2) and index._identifiers.shape == (5,)
    # recall over that index: true ids chosen so that hits@1 = 3, hits@2 = 4
    metric = IndexRecall(index, ks=[1, 2])
    truth = np.array(["candidate_1", "candidate_4", "candidate_5", "candidate_9", "candidate_1"], dtype=object).reshape(5, 1)
    metric(inputs, truth)
    assert metric.hits[1] == 3 and metric.hits[2] == 4 and metric.metric[2] == np.float64(0.8)
    with pytest.raises(ValueError):
        BruteForceIndex(6, qm, pairs)


def _c2_features(vq, vc):
    qf = [Feature("age", tt.float32, FeatureFamily.QUERY),
          Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=64)]
    cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=64),
          Feature("product_type_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=16),
          Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=8)]
    qf[1].set_vocab_size(vq); cf[0].set_vocab_size(vc); cf[1].set_vocab_size(131); cf[2].set_vocab_size(50)
    return qf, cf


def _oracle_twin(m):
    def tower(t):
        feats = [O.OracleFeature(f.name, f.dtype == tt.string, f.embedding_size) for f in t.features]
        tables = {n: _np(e.weight).copy() for n, e in t.input_layer.embedding_layers.items()}
        dense = [(_np(w).copy(), _np(b).copy()) for w, b in zip(t.kernels, t.biases)]
        return O.OracleTower(feats, tables, dense)
    return tower(m.query_tower), tower(m.candidate_tower)


def _batch(rng, B, vq, vc):
    art = np.minimum(rng.zipf(1.2, size=B), vc).astype(np.int32)            # duplicates in the batch
    return {"age": rng.random((B, 1)).astype(np.float32), "customer_id": rng.integers(0, vq + 1, size=(B, 1)).astype(np.int32),
            "article_id": art.reshape(B, 1), "product_type_name": (art % 131 + 1).reshape(B, 1).astype(np.int32),
            "colour_group_name": (art % 50 + 1).reshape(B, 1).astype(np.int32)}


@pytest.mark.parametrize("hidden", [None, [96]])
@pytest.mark.parametrize("impl", [1])
def test_train_step_matches_oracle_c2_shape(lib, hidden, impl):
    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    set_seed(99)
    vq, vc, B = 5000, 800, 777
    qf, cf = _c2_features(vq, vc)
    rng = np.random.default_rng(21)
    probs = {str(i + 1): float(p) for i, p in enumerate(rng.dirichlet(np.ones(vc)))}
    m = TwoTowerModel(qf, cf, "article_id", 64, hidden, hidden, candidate_prob_lookup=probs)
    m.impl = impl
    m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
    qt, ct = _oracle_twin(m)
    data = _batch(rng, B, vq, vc)
    p_rows = np.ones(vc + 1, np.float32); p_rows[1:] = [np.float32(probs[str(i + 1)]) for i in range(vc)]
    col_p = p_rows[data["article_id"].reshape(-1)]
    ids_q = {"customer_id": data["customer_id"]}; ids_c = {k: data[k] for k in ("article_id", "product_type_name", "colour_group_name")}
    g = O.train_step_grads(qt, ct, ids_q, {"age": data["age"]}, ids_c, {}, col_p)
    logits = _np(m(data))
    # north star: logits within 1e-3 relative (fp32 exact path: bit-level agreement up to fp64-vs-fp32 accumulation)
    np.testing.assert_allclose(logits, g.logits + np.log(col_p)[None, :], rtol=1e-5, atol=1e-6)
    out = m.train_step(data)
    assert abs(float(out["loss"]) - g.loss) <= 1e-5 * abs(g.loss)
    # apply the oracle's optimizer and compare every parameter
    for tower, ot, dg, sl in ((m.query_tower, qt, g.dense_q, g.tables_q), (m.candidate_tower, ct, g.dense_c, g.tables_c)):
        for i, (w, b) in enumerate(ot.dense):
            aw, ab = np.full_like(w, 0.1), np.full_like(b, 0.1)
            O.adagrad_dense(w, aw, dg[i][0], 0.05); O.adagrad_dense(b, ab, dg[i][1], 0.05)
            # Adagrad's first step moves each weight by ~lr*sign(g): compare the update, not the weight
            np.testing.assert_allclose(_np(tower.kernels[i]), w, rtol=0, atol=2e-4)
            np.testing.assert_allclose(_np(tower.biases[i]), b, rtol=0, atol=2e-4)
        for name, s in sl.items():
            t = ot.tables[name]; acc = np.full_like(t, 0.1)
            O.adagrad_sparse(t, acc, s, 0.05)
            got = _np(tower.input_layer.embedding_layers[name].weight)
            np.testing.assert_allclose(got, t, rtol=0, atol=2e-4)
            untouched = np.setdiff1d(np.arange(t.shape[0]), s.indices)
            assert np.array_equal(got[untouched], t[untouched])                # only touched rows move


def test_training_is_deterministic_and_graph_equals_eager(lib):
    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    def run(graph):
        set_seed(5)
        qf, cf = _c2_features(3000, 500)
        m = TwoTowerModel(qf, cf, "article_id", 64, candidate_prob_lookup={str(i + 1): 1.0 / 500 for i in range(500)})
        m.use_cuda_graph = graph
        m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
        rng = np.random.default_rng(1)
        losses = []
        for _ in range(5):
            losses.append(float(m.train_step(_batch(rng, 512, 3000, 500))["loss"]))
        state = m.state_arrays()
        return losses, state

    l1, s1 = run(False)
    l2, s2 = run(False)
    l3, s3 = run(True)
    assert l1 == l2 and all(np.array_equal(s1[k], s2[k]) for k in s1)          # run-to-run bit-identical
    assert l1 == l3 and all(np.array_equal(s1[k], s3[k]) for k in s1)          # CUDA-graph replay == eager
    assert l1[-1] < l1[0]                                                        # and it learns


def test_fit_accepts_ragged_final_batch_and_save(lib, tmp_path):
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    qf, cf = _c2_features(1000, 300)
    m = TwoTowerModel(qf, cf, "article_id", 32)
    m.compile(optimizer=OptimizerFactory.get_optimizer("adam", {"learning_rate": 0.001}))
    rng = np.random.default_rng(2)
    ds = [_batch(rng, b, 1000, 300) for b in (256, 256, 100)]                    # drop_remainder=False (tfrecord_dataset.py:97)
    h = m.fit(ds, epochs=2, verbose=0)
    assert len(h["loss"]) == 2 and np.isfinite(h["loss"]).all() and h["loss"][1] < h["loss"][0]
    m.save(str(tmp_path / "model") + "/")
    for sub in ("two_tower", "query_tower", "candidate_tower"):                  # two_tower_model.py:176-205: dirname(model_path)/<sub>
        assert (tmp_path / "model" / sub / "variables.npz").exists()


def test_adam_train_step_matches_oracle(lib):
    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    set_seed(7)
    qf, cf = _c2_features(400, 120)
    m = TwoTowerModel(qf, cf, "article_id", 32)
    m.impl = 1   # exact fp32 contraction: Adam's first step is ~lr*sign(g), so TF32 noise would flip signs where g ~ 0
    m.compile(optimizer=OptimizerFactory.get_optimizer("adam", {"learning_rate": 0.01}))
    qt, ct = _oracle_twin(m)
    rng = np.random.default_rng(3)
    data = _batch(rng, 200, 400, 120)
    g = O.train_step_grads(qt, ct, {"customer_id": data["customer_id"]}, {"age": data["age"]},
                           {k: data[k] for k in ("article_id", "product_type_name", "colour_group_name")}, {}, None)
    m.train_step(data)
    t = ct.tables["article_id"]; mm = np.zeros_like(t); vv = np.zeros_like(t)
    O.adam_sparse(t, mm, vv, g.tables_c["article_id"], 0.01, 1)
    got = _np(m.candidate_tower.input_layer.embedding_layers["article_id"].weight)
    np.testing.assert_allclose(got, t, rtol=0, atol=5e-4)   # first Adam step is ~lr*sign(g); sign flips only where |g|~0
    w, b = ct.dense[0]; mw = np.zeros_like(w); vw = np.zeros_like(w)
    O.adam_dense(w, mw, vw, g.dense_c[0][0], 0.01, 1)
    np.testing.assert_allclose(_np(m.candidate_tower.kernels[0]), w, rtol=0, atol=5e-4)


def test_index_over_trained_towers_recall_at_12_bit_exact(lib):
    """C2-style eval: candidate tower -> BruteForceIndex -> Recall@12, indices bit-exact vs the oracle."""
    from pkg.modelling._device import set_seed
    from pkg.modelling.indices.brute_force import BruteForceIndex
    from pkg.modelling.metrics.index_recall import IndexRecall
    from pkg.modelling.models.two_tower_model import TwoTowerModel

    set_seed(11)
    vq, vc = 2000, 3000
    qf, cf = _c2_features(vq, vc)
    m = TwoTowerModel(qf, cf, "article_id", 64)
    art = np.arange(1, vc + 1, dtype=np.int32)
    pairs = []
    for lo in range(0, vc, 1000):                                                # candidate_batch_size
        a = art[lo:lo + 1000]
        x = {"article_id": a.reshape(-1, 1), "product_type_name": (a % 131 + 1).reshape(-1, 1), "colour_group_name": (a % 50 + 1).reshape(-1, 1)}
        pairs.append((a, m.candidate_tower(x)))
    index = BruteForceIndex(12, m.query_tower, pairs)
    index.impl = 1
    rng = np.random.default_rng(4)
    B = 300
    queries = {"age": rng.random((B, 1)).astype(np.float32), "customer_id": rng.integers(0, vq + 1, size=(B, 1)).astype(np.int32)}
    truth = rng.integers(1, vc + 1, size=(B, 1)).astype(np.int32)
    got = index(queries)
    q_emb = _np(m.query_tower(queries)); c_emb = _np(index._candidates)
    _, want = O.index_topk(q_emb, c_emb, 12)
    assert np.array_equal(got, art[want])
    metric = IndexRecall(index, [1, 12])
    metric(queries, truth)
    rec = O.RecallOracle([1, 12]); rec.update(truth, art[want])
    assert metric.hits[12] == rec.hits[12] and metric.metric[12] == rec.metric[12]


def test_save_load_round_trip_and_bulk_index_build(lib, tmp_path):
    """weights survive save -> load into a fresh model (bit-exact tower outputs); an index built by
    BruteForceIndex.from_candidate_tower equals the one built from per-batch (ids, embeddings) pairs; a saved index reloads."""
    import torch

    from pkg.modelling._device import set_seed
    from pkg.modelling.indices.brute_force import BruteForceIndex
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory
    from pkg.schema import dtypes as tt
    from pkg.schema.features import Feature, FeatureFamily

    def make(seed):
        set_seed(seed)
        qf = [Feature("age", tt.float32, FeatureFamily.QUERY), Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=32)]
        cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=32),
              Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=8)]
        qf[1].set_vocab_size(500); cf[0].set_vocab_size(6000); cf[1].set_vocab_size(50)
        m = TwoTowerModel(qf, cf, "article_id", 32, query_tower_units=[48])
        m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
        return m

    rng = np.random.default_rng(3)
    m = make(1)
    art = rng.integers(1, 6001, size=(256, 1)).astype(np.int32)
    batch = {"age": rng.random((256, 1)).astype(np.float32), "customer_id": rng.integers(0, 501, size=(256, 1)).astype(np.int32),
             "article_id": art, "colour_group_name": (art % 50 + 1).astype(np.int32)}
    m.train_step(batch)
    m.save(str(tmp_path / "model") + "/")
    m2 = make(2)                                              # different initialisation
    q = {"age": batch["age"], "customer_id": batch["customer_id"]}
    assert not torch.equal(m.query_tower(q), m2.query_tower(q))
    m2.load(str(tmp_path / "model") + "/")
    assert torch.equal(m.query_tower(q), m2.query_tower(q))
    c = {"article_id": batch["article_id"], "colour_group_name": batch["colour_group_name"]}
    assert torch.equal(m.candidate_tower(c), m2.candidate_tower(c))
    m3 = make(3)
    m3.candidate_tower.load(str(tmp_path / "model" / "candidate_tower"))   # single-tower artefact
    assert torch.equal(m.candidate_tower(c), m3.candidate_tower(c))
    # bulk build == per-batch pairs (runner.py:88-93)
    ids = np.arange(1, 6001, dtype=np.int32)
    cand_batches = [{"article_id": ids[lo:lo + 2500].reshape(-1, 1), "colour_group_name": (ids[lo:lo + 2500] % 50 + 1).reshape(-1, 1)}
                    for lo in range(0, 6000, 2500)]
    pairs = [(b["article_id"].reshape(-1), m.candidate_tower(b)) for b in cand_batches]
    a = BruteForceIndex(12, m.query_tower, pairs)
    b = BruteForceIndex.from_candidate_tower(12, m.query_tower, m.candidate_tower, cand_batches, "article_id")
    assert torch.equal(a._candidates, b._candidates) and np.array_equal(a._identifiers, b._identifiers)
    want = a(q)
    assert np.array_equal(want, b(q))
    b.save(str(tmp_path / "index"))
    c2 = BruteForceIndex.load(str(tmp_path / "index"), 12, m.query_tower)
    assert np.array_equal(want, c2(q))


def test_c1_config_full_vocabularies_train_step_and_top12_index(lib):
    """BASELINE configs[0] at its real sizes: id-only towers (1 371 980 customers + OOV, 105 542 articles + OOV), embedding = joint
    = 32, batch 1024, logQ in-batch softmax, Adagrad; then the brute-force top-12 index over all 105 542 candidate-tower rows."""
    from pkg.modelling._device import set_seed
    from pkg.modelling.indices.brute_force import BruteForceIndex
    from pkg.modelling.metrics.index_recall import IndexRecall
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    set_seed(5)
    vq, vc, B = 1_371_980, 105_542, 1024
    qf = [Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=32)]
    cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=32)]
    qf[0].set_vocab_size(vq); cf[0].set_vocab_size(vc)
    rng = np.random.default_rng(31)
    w = 1.0 / np.arange(1, vc + 1)
    p = (w / w.sum()).astype(np.float32)                                       # Zipf(1) sampling probabilities
    probs = {str(i + 1): float(p[i]) for i in range(vc)}
    m = TwoTowerModel(qf, cf, "article_id", 32, candidate_prob_lookup=probs)
    m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
    qt, ct = _oracle_twin(m)
    art = (np.searchsorted(np.cumsum(p.astype(np.float64)), rng.random(B)) + 1).clip(1, vc).astype(np.int32)
    data = {"customer_id": rng.integers(1, vq + 1, size=(B, 1)).astype(np.int32), "article_id": art.reshape(B, 1)}
    p_rows = np.ones(vc + 1, np.float32); p_rows[1:] = p
    g = O.train_step_grads(qt, ct, {"customer_id": data["customer_id"]}, {}, {"article_id": data["article_id"]}, {}, p_rows[art])
    loss = float(m.train_step(data)["loss"])
    assert abs(loss - g.loss) <= 1e-3 * abs(g.loss)                           # north star (TF32 tensor-core path at E = 32)
    for tower, ot, sl in ((m.query_tower, qt, g.tables_q), (m.candidate_tower, ct, g.tables_c)):
        for name, s in sl.items():
            t = ot.tables[name]; acc = np.full_like(t, 0.1)
            O.adagrad_sparse(t, acc, s, 0.05)
            got = _np(tower.input_layer.embedding_layers[name].weight)
            rows = np.unique(s.indices)
            np.testing.assert_allclose(got[rows], t[rows], rtol=0, atol=2e-3)
            probe = rng.integers(0, t.shape[0], 4096)
            probe = probe[~np.isin(probe, rows)]
            assert np.array_equal(got[probe], t[probe])                          # untouched rows do not move
    # index over the whole candidate vocabulary, top-12, bit-exact against the canonical oracle
    ids = np.arange(1, vc + 1, dtype=np.int32)
    index = BruteForceIndex.from_candidate_tower(12, m.query_tower, m.candidate_tower,
                                                 [{"article_id": ids[lo:lo + 10000].reshape(-1, 1)} for lo in range(0, vc, 10000)], "article_id")
    q = {"customer_id": rng.integers(1, vq + 1, size=(256, 1)).astype(np.int32)}
    got = index(q)
    q_emb = _np(m.query_tower(q)); c_emb = _np(index._candidates)
    _, want = O.index_topk(q_emb, c_emb, 12)
    assert np.array_equal(got, ids[want])
    metric = IndexRecall(index, ks=[1, 12])
    metric(q, ids[want[:, 3]].reshape(-1, 1))                                   # "truth" = each query's 4th-ranked article
    assert metric.hits[1] == 0 and metric.hits[12] == 256 and metric.metric[12] == np.float64(1.0)


def test_recompile_after_training_uses_the_new_optimizer_slots(lib):
    """compile() / load() after training replace the optimizer slot tensors; the cached step workspaces (job lists, captured
    graphs) pointed at the old ones.  A recompiled model must continue exactly like a fresh model holding the same weights."""
    import torch

    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    def fresh():
        set_seed(11)
        qf, cf = _c2_features(3000, 500)
        m = TwoTowerModel(qf, cf, "article_id", 64, candidate_prob_lookup={str(i + 1): 1.0 / 500 for i in range(500)})
        m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
        return m

    rng = np.random.default_rng(3)
    b1, b2 = _batch(rng, 512, 3000, 500), _batch(rng, 512, 3000, 500)
    a = fresh()
    a.train_step(b1)
    a.train_step(b1)                                   # second call replays the captured graph
    weights = {k: v.copy() for k, v in a.state_arrays().items()}
    a.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))     # fresh accumulators (0.1)
    junk = [torch.full((1 << 20,), 7.0, device="cuda") for _ in range(8)]                        # recycle the freed slot memory
    la = float(a.train_step(b2)["loss"])
    b = fresh()
    b.query_tower.load_state_arrays(weights, "query_tower/")
    b.candidate_tower.load_state_arrays(weights, "candidate_tower/")
    sb0 = b.state_arrays()
    assert all(np.array_equal(sb0[k], weights[k]) for k in weights)
    lb = float(b.train_step(b2)["loss"])
    sa, sb = a.state_arrays(), b.state_arrays()
    assert la == lb and all(np.array_equal(sa[k], sb[k]) for k in sa)
    del junk


def test_embedding_rows_do_not_depend_on_the_sharding(lib):
    """tt_fill_uniform: a cell is a function of (seed, global row, column), so shards filled by their owners hold what the whole
    table holds; values are uniform on [-0.05, 0.05) like tf-keras' Embedding default."""
    import torch

    from pkg import _native as N

    rows, e, seed = 10_007, 24, 1234
    full = torch.empty((rows, e), device="cuda")
    N.check(lib.tt_fill_uniform(full.data_ptr(), rows, e, 0, 1, seed, -0.05, 0.05, N.stream_ptr()))
    for world in (2, 3, 8):
        for r in range(world):
            n_r = (rows - r + world - 1) // world
            part = torch.full((n_r + 1, e), 9.0, device="cuda")
            N.check(lib.tt_fill_uniform(part.data_ptr(), n_r, e, r, world, seed, -0.05, 0.05, N.stream_ptr()))
            assert torch.equal(part[:n_r], full[r::world]) and bool((part[n_r] == 9.0).all())
    f = full.double()
    assert float(f.min()) >= -0.05 and float(f.max()) < 0.05
    assert abs(float(f.mean())) < 5e-4 and abs(float(f.std()) - 0.1 / np.sqrt(12)) < 5e-4
    assert abs(float(torch.corrcoef(torch.stack([f[:-1, 0], f[1:, 0]]))[0, 1])) < 0.05      # neighbouring rows uncorrelated
    other = torch.empty_like(full)
    N.check(lib.tt_fill_uniform(other.data_ptr(), rows, e, 0, 1, seed + 1, -0.05, 0.05, N.stream_ptr()))
    assert not torch.equal(other, full)
    assert lib.tt_fill_uniform(None, 4, e, 0, 1, seed, -0.05, 0.05, None) != 0                # null pointer is an error, not a crash


def test_index_from_local_rows_equals_the_whole_corpus_index(lib):
    """from_local_rows: two shards built from their own rows, merged, answer like the index over the whole corpus; without
    identifiers the call returns global row numbers."""
    import torch

    from pkg import _native as N
    from pkg.modelling.indices.brute_force import BruteForceIndex

    n, e, k, nq = 20_000, 64, 50, 96
    g = torch.Generator(device="cuda").manual_seed(8)
    corpus = torch.randn((n, e), generator=g, device="cuda") * 0.3
    q = torch.randn((nq, e), generator=g, device="cuda")
    class Ident:                    # query "tower": the embeddings themselves
        def __call__(self, x):
            return x["q"]

        def get_input_signature(self):
            return {}

    ident = Ident()
    whole = BruteForceIndex(k, ident, [(np.arange(n, dtype=np.int32), corpus)])
    s0, i0 = whole.search(q)
    cut = 7_777
    parts = [BruteForceIndex.from_local_rows(k, ident, corpus[:cut], 0, n), BruteForceIndex.from_local_rows(k, ident, corpus[cut:], cut, n)]
    ss, ii = zip(*(p.search(q) for p in parts))
    all_s, all_i = torch.stack(ss).contiguous(), torch.stack(ii).contiguous()
    out_s, out_i = torch.empty_like(s0), torch.empty_like(i0)
    N.check(lib.tt_topk_merge(all_s.data_ptr(), all_i.data_ptr(), 2, nq, k, out_s.data_ptr(), out_i.data_ptr(), N.stream_ptr()))
    assert torch.equal(out_i, i0) and torch.equal(out_s, s0)
    one = BruteForceIndex.from_local_rows(k, ident, corpus, 0, n)
    got = one({"q": q})
    assert got.dtype == np.int32 and np.array_equal(got, i0.cpu().numpy())
    pinned = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    view = one({"q": q}, out=pinned)
    assert np.array_equal(view, got) and view.ctypes.data == pinned.data_ptr()
    assert np.array_equal(one.positions_of([0, 5, n - 1, n, -3]), [0, 5, n - 1, -1, -1])
    with pytest.raises(ValueError):
        BruteForceIndex.from_local_rows(k, ident, corpus[:10], n - 5, n)


def test_recall_counts_every_row_of_a_repeated_identifier(lib):
    """The reference compares identifiers (tf.equal): when an id occurs twice in the corpus, a hit on EITHER row counts (and both
    count when both are within the cut-off).  The device path compares row indices, so it maps rows to the id's first row."""
    import torch

    from pkg.modelling.indices.brute_force import BruteForceIndex
    from pkg.modelling.metrics.index_recall import IndexRecall

    class Ident:
        def __call__(self, x):
            return x["q"]

        def get_input_signature(self):
            return {}

    rng = np.random.default_rng(4)
    n, e = 300, 16
    corpus = rng.standard_normal((n, e)).astype(np.float32)
    ids = np.array([f"id{i}" for i in range(n)], dtype=object)
    ids[200:220] = ids[0:20]                                   # twenty identifiers occur twice
    corpus[200:220] = corpus[0:20] * 1.5                       # ... and the SECOND row scores higher for queries along that row
    index = BruteForceIndex(5, Ident(), [(ids, corpus)])
    assert index.canonical_rows() is not None
    q = corpus[:40].copy()                                     # query b is most similar to rows b (and 200 + b for b < 20)
    truth = ids[:40].reshape(-1, 1)
    got = IndexRecall(index, ks=[1, 5])({"q": torch.from_numpy(q).cuda()}, truth)
    cand = index({"q": torch.from_numpy(q).cuda()})            # identifiers, as the reference compares them
    tb = np.array([t.encode() if isinstance(t, str) else t for t in truth.reshape(-1)]).reshape(-1, 1)
    cb = np.array([[c.encode() if isinstance(c, str) else c for c in row] for row in cand])
    want = {k: float(np.sum(tb == cb[:, :k])) / 40 for k in (1, 5)}
    assert {k: float(v) for k, v in got.items()} == want
    assert want[5] > 1.0 - 1e-9                                # the twenty duplicated ids are found twice within the top 5
