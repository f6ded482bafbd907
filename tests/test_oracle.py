"""The oracle against the reference's own fixtures and the known-answer vectors (CPU only)."""
import numpy as np
import pytest

from oracle import two_tower_oracle as O


def test_logq_matches_reference_fixture(golden):
    fix, _ = golden
    g = fix["logq"]
    probs = O.prob_lookup(g["candidate_prob_lookup"], np.array(g["candidate_ids"]).reshape(-1, 1))
    z = O.logq_correction(np.array(g["logits"], np.float32), probs)
    # tests/test_layers.py:28-36; the literals are correct to 1e-10, fp32 leaves ~2e-7
    np.testing.assert_allclose(z, np.array(g["expected"]), rtol=0, atol=5e-7)
    # unknown id -> default probability 1.0 -> no correction (logq_correction.py:38-41)
    assert O.prob_lookup(g["candidate_prob_lookup"], ["nope"])[0] == np.float32(1.0)


@pytest.mark.parametrize("canonical", [False, True])
def test_brute_force_matches_reference_fixture(golden, canonical):
    fix, _ = golden
    g = fix["brute_force"]
    rows = O.string_lookup(g["query_vocab"], np.array(g["queries"]).reshape(-1, 1)).reshape(-1)
    assert rows.tolist() == [1, 2, 3, 0, 1]          # query_4 is OOV -> row 0
    q = np.array(g["query_table"], np.float32)[rows]
    c = np.array(g["candidate_embeddings"], np.float32)
    if canonical:
        _, idx = O.index_topk(q, c, g["k"])
    else:
        _, idx = O.top_k(O.logits_qct(q, c), g["k"])
    got = np.array(g["candidate_ids"])[idx]
    assert got.tolist() == g["expected"]


def test_recall_matches_reference_fixture(golden):
    fix, _ = golden
    g = fix["recall"]
    static = np.array(g["static_candidates"]).reshape(1, -1)
    rec = O.RecallOracle(g["ks"])
    truth = np.array(g["true_candidate_ids"])
    bs = g["batch_size"]
    for i in range(0, len(truth), bs):                 # ragged final batch (2, 2, 1)
        t = truth[i:i + bs]
        rec.update(t, O.static_index_call(static, g["static_k"], len(t)))
    for k, v in g["expected"].items():
        assert rec.metric[int(k)] == np.float64(v)
        assert rec.metric[int(k)].dtype == np.float64
    assert rec.hits[1] == 1 and rec.hits[2] == 3 and rec.seen == 5


def _kat_towers(a):
    qt = O.OracleTower([O.OracleFeature("q", True, 2)], {"q": np.array(a["Tq"], np.float32)},
                       [(np.array(a["Wq"], np.float32), np.array(a["bq"], np.float32))])
    ct = O.OracleTower([O.OracleFeature("c", True, 2)], {"c": np.array(a["Tc"], np.float32)},
                       [(np.array(a["Wc"], np.float32), np.array(a["bc"], np.float32))])
    return qt, ct


def test_kat_a_full_train_step(golden):
    _, kat = golden
    a = kat["A"]
    qt, ct = _kat_towers(a)
    qid, cid = np.array(a["query_ids"]), np.array(a["candidate_ids"])
    probs = np.array(a["p_row"], np.float32)[cid]
    g = O.train_step_grads(qt, ct, {"q": qid}, {}, {"c": cid}, {}, probs)
    assert abs(g.loss - a["loss"]) < 1e-5
    np.testing.assert_allclose(g.q, a["Q"], atol=1e-7)
    np.testing.assert_allclose(g.c, a["C"], atol=1e-7)
    np.testing.assert_allclose(g.logits, a["Z"], atol=1e-6)
    np.testing.assert_allclose(g.dq, a["dQ"], atol=1e-6)
    np.testing.assert_allclose(g.dc, a["dC"], atol=1e-6)
    np.testing.assert_allclose(g.dense_q[0][0], a["dWq"], atol=1e-6)
    np.testing.assert_allclose(g.dense_q[0][1], a["dbq"], atol=1e-6)
    np.testing.assert_allclose(g.dense_c[0][0], a["dWc"], atol=1e-6)
    np.testing.assert_allclose(g.dense_c[0][1], a["dbc"], atol=1e-6)
    np.testing.assert_allclose(g.tables_c["c"].values, a["dxc"], atol=1e-6)
    # sparse Adagrad: duplicates summed FIRST (row 1 appears twice)
    tc = np.array(a["Tc"], np.float32); acc = np.full_like(tc, a["acc0"])
    O.adagrad_sparse(tc, acc, g.tables_c["c"], a["lr"])
    np.testing.assert_allclose(tc, a["Tc_after"], atol=1e-6)
    np.testing.assert_allclose(acc, a["acc_c_after"], atol=1e-6)
    assert not np.allclose(tc[1], [0.54506421542, 0.52387219416], atol=1e-4)  # per-occurrence accumulation must NOT match
    tq = np.array(a["Tq"], np.float32); accq = np.full_like(tq, a["acc0"])
    O.adagrad_sparse(tq, accq, g.tables_q["q"], a["lr"])
    np.testing.assert_allclose(tq, a["Tq_after"], atol=1e-6)
    w = np.array(a["Wc"], np.float32); wacc = np.full_like(w, a["acc0"])
    O.adagrad_dense(w, wacc, g.dense_c[0][0], a["lr"])
    np.testing.assert_allclose(w, a["Wc_after"], atol=1e-6)


def test_kat_b_ce_sum_on_reference_logq_fixture(golden):
    fix, kat = golden
    g = fix["logq"]
    probs = O.prob_lookup(g["candidate_prob_lookup"], g["candidate_ids"])
    z = O.logq_correction(np.array(g["logits"], np.float32), probs)
    loss, _, dz = O.ce_sum_from_logits(z)
    assert abs(loss - kat["B"]["loss"]) < 2e-6
    np.testing.assert_allclose(dz.sum(axis=1), 0.0, atol=1e-12)   # softmax - onehot sums to 0 per row


def test_kat_c_topk_tie_break(golden):
    _, kat = golden
    c = kat["C"]
    s = np.array(c["scores"], np.float32)
    _, i1 = O.top_k(s, c["k"])
    _, i2 = O.top_k_numpy(s, c["k"])
    assert i1.tolist() == c["indices"] and i2.tolist() == c["indices"]


def test_topk_heap_equals_stable_sort_with_many_ties():
    rng = np.random.default_rng(7)
    s = rng.integers(0, 5, size=(33, 257)).astype(np.float32)     # heavy ties
    for k in (1, 7, 100, 257):
        s1, i1 = O.top_k(s, k)
        s2, i2 = O.top_k_numpy(s, k)
        assert np.array_equal(i1, i2) and np.array_equal(s1, s2)


def test_canonical_scores_are_sequential_fmaf_and_dyadic_exact():
    rng = np.random.default_rng(3)
    # dyadic grid: every partial sum is exact, so every evaluation order agrees bit for bit
    q = (rng.integers(-16, 17, size=(9, 64)) / 16.0).astype(np.float32)
    c = (rng.integers(-16, 17, size=(50, 64)) / 16.0).astype(np.float32)
    assert np.array_equal(O.logits_qct(q, c, canonical=True), O.logits_qct(q, c, canonical=False))
    # generic inputs: canonical == explicit python loop with exact products in float64 rounded per step
    q = rng.standard_normal((3, 16)).astype(np.float32); c = rng.standard_normal((4, 16)).astype(np.float32)
    got = O.logits_qct(q, c, canonical=True)
    for i in range(3):
        for j in range(4):
            acc = np.float32(0)
            for k in range(16):
                # fma: exact product (fits float64) + acc, single rounding to fp32 (float64 sum is exact enough
                # here: |acc| and product differ by < 2^29, so the float64 sum is exact)
                acc = np.float32(np.float64(q[i, k]) * np.float64(c[j, k]) + np.float64(acc))
            assert acc == got[i, j]


def test_dedup_sum_is_position_ordered_fp32():
    vals = np.array([[1e8], [1.0], [-1e8], [1.0]], np.float32)
    s = O.dedup_indexed_slices(O.IndexedSlices(np.array([5, 5, 5, 5]), vals))
    # ((1e8 + 1) - 1e8) + 1 in fp32 == 1 (the first +1 is absorbed); any other order gives 2 or 0
    assert s.indices.tolist() == [5] and s.values[0, 0] == np.float32(1.0)


def test_merge_topk_is_independent_of_sharding():
    rng = np.random.default_rng(11)
    q = rng.integers(0, 4, size=(6, 8)).astype(np.float32)
    c = rng.integers(0, 4, size=(120, 8)).astype(np.float32)
    s_all, i_all = O.index_topk(q, c, 10)
    for g in (2, 3, 8):
        per = (120 + g - 1) // g
        parts = [O.index_topk(q, c[r * per:(r + 1) * per], 10, idx_base=r * per) for r in range(g)]
        s, i = O.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10)
        assert np.array_equal(i, i_all) and np.array_equal(s, s_all)


def test_adam_sparse_is_not_lazy():
    t = np.ones((4, 2), np.float32); m = np.full((4, 2), 0.5, np.float32); v = np.full((4, 2), 0.25, np.float32)
    O.adam_sparse(t, m, v, O.IndexedSlices(np.array([2]), np.array([[1.0, -1.0]], np.float32)), lr=0.1, step=1)
    assert np.all(t[0] != 1.0) and np.all(m[0] == np.float32(0.5) * np.float32(0.9))   # untouched rows still decay and move
