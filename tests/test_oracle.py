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


def _c2_like_problem(seed=7, B=24, E=8):
    """C2-shaped towers: numeric feature, side features, a hidden layer, duplicate ids in the batch, logQ probabilities."""
    rng = np.random.default_rng(seed)

    def tower(feats, rows, hidden):
        tables = {f.name: rng.uniform(-0.5, 0.5, (rows[f.name], f.embedding_size)).astype(np.float32) for f in feats if f.is_string}
        d = sum(1 if not f.is_string else f.embedding_size for f in feats)
        dims = [d] + hidden + [E]
        dense = [(rng.uniform(-0.6, 0.6, (dims[i], dims[i + 1])).astype(np.float32), rng.uniform(0.05, 0.3, dims[i + 1]).astype(np.float32))
                 for i in range(len(dims) - 1)]
        return O.OracleTower(feats, tables, dense)

    qt = tower([O.OracleFeature("age", False), O.OracleFeature("q", True, 6)], {"q": 40}, [10])
    ct = tower([O.OracleFeature("c", True, 6), O.OracleFeature("colour", True, 3)], {"c": 12, "colour": 5}, [])
    q_ids = {"q": rng.integers(0, 40, B)}
    q_num = {"age": rng.uniform(0.1, 0.9, B).astype(np.float32)}
    c_ids = {"c": rng.integers(0, 12, B), "colour": rng.integers(0, 5, B)}          # 24 draws from 12 rows: duplicates
    probs = rng.uniform(0.01, 0.3, B).astype(np.float32)
    return rng, qt, ct, q_ids, q_num, c_ids, probs


def _dense_table_grads(t, slices):
    out = {}
    for k in sorted(t.tables):
        d = np.zeros(t.tables[k].shape, np.float64)
        np.add.at(d, slices[k].indices, slices[k].values.astype(np.float64))
        out[k] = d
    return out


def test_gradients_equal_torch_autograd_of_the_forward():
    """The reference differentiates its forward with tf.GradientTape (two_tower_model.py:110-124).  An independent float64
    torch forward (gather, concat numerics-first, relu Dense on every layer, Q.C^T - ln p, CE SUM against eye) differentiated by
    autograd must give the oracle's hand-written backward."""
    import torch

    _, qt, ct, q_ids, q_num, c_ids, probs = _c2_like_problem()

    def leaves(t):
        tabs = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in t.tables.items()}
        dense = [(torch.tensor(w, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True))
                 for w, b in t.dense]
        return tabs, dense

    def fwd(t, tabs, dense, ids, nums):
        cols = [torch.tensor(nums[f.name], dtype=torch.float64).reshape(-1, 1) for f in t.numerical]
        cols += [tabs[f.name][torch.tensor(ids[f.name])] for f in t.categorical]
        h = torch.cat(cols, dim=1)
        for w, b in dense:
            h = torch.relu(h @ w + b)
        return h

    qtab, qden = leaves(qt)
    ctab, cden = leaves(ct)
    z = fwd(qt, qtab, qden, q_ids, q_num) @ fwd(ct, ctab, cden, c_ids, {}).T - torch.log(torch.tensor(probs, dtype=torch.float64))[None, :]
    loss = torch.nn.functional.cross_entropy(z, torch.arange(z.shape[0]), reduction="sum")
    loss.backward()
    g = O.train_step_grads(qt, ct, q_ids, q_num, c_ids, {}, probs)
    assert abs(g.loss - loss.item()) < 1e-4 * loss.item()
    for t, tabs, den, sl, dg in ((qt, qtab, qden, g.tables_q, g.dense_q), (ct, ctab, cden, g.tables_c, g.dense_c)):
        for k, d in _dense_table_grads(t, sl).items():
            np.testing.assert_allclose(d, tabs[k].grad.numpy(), rtol=2e-4, atol=2e-5)
        for (w, b), (dw, db) in zip(den, dg):
            np.testing.assert_allclose(dw, w.grad.numpy(), rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(db, b.grad.numpy(), rtol=2e-4, atol=2e-5)


def test_gradients_equal_finite_differences_of_the_loss():
    """No reference test pins a gradient, so the oracle's backward pass is also checked against the derivative
    of its own forward loss: directional central differences, one random direction per parameter tensor."""
    rng, qt, ct, q_ids, q_num, c_ids, probs = _c2_like_problem()

    def params(t):
        return [t.tables[k] for k in sorted(t.tables)] + [a for wb in t.dense for a in wb]

    def loss():
        return O.train_step_grads(qt, ct, q_ids, q_num, c_ids, {}, probs).loss

    g = O.train_step_grads(qt, ct, q_ids, q_num, c_ids, {}, probs)

    def grads(t, slices, dense):
        out = []
        for k in sorted(t.tables):
            d = np.zeros(t.tables[k].shape, np.float64)
            np.add.at(d, slices[k].indices, slices[k].values.astype(np.float64))
            out.append(d)
        return out + [a.astype(np.float64) for wb in dense for a in wb]

    ps = params(qt) + params(ct)
    gs = grads(qt, g.tables_q, g.dense_q) + grads(ct, g.tables_c, g.dense_c)
    assert [p.shape for p in ps] == [x.shape for x in gs]
    # one random direction per parameter tensor; h small enough that ReLU kinks crossed by the step do not matter
    # (observed agreement 1e-3 .. 3e-3; a wrong formula is off by O(1))
    h = 2e-4
    for p, x in zip(ps, gs):
        for _ in range(2):
            d = rng.standard_normal(p.shape)
            keep = p.copy()
            p[...] = (keep.astype(np.float64) + h * d).astype(np.float32)
            up = loss()
            p[...] = (keep.astype(np.float64) - h * d).astype(np.float32)
            down = loss()
            p[...] = keep
            numeric, analytic = (up - down) / (2 * h), float(np.sum(x * d))
            assert abs(numeric - analytic) <= 1e-2 * max(abs(analytic), 1.0), (p.shape, numeric, analytic)


def test_adagrad_equals_an_independent_implementation():
    """tf-keras legacy Adagrad (acc0 = 0.1, eps = 1e-7 added OUTSIDE the square root) is the rule torch.optim.Adagrad implements;
    the sparse path with duplicates summed first must equal the dense rule applied to the densified gradient (untouched rows have
    zero gradient and do not move).  Four steps, fp32."""
    import torch

    rng = np.random.default_rng(11)
    rows, e, lr = 9, 4, 0.05
    table = rng.uniform(-0.05, 0.05, (rows, e)).astype(np.float32)
    acc = np.full_like(table, O.ADAGRAD_INIT_ACC)
    dense_w, dense_acc = table.copy(), acc.copy()
    p = torch.nn.Parameter(torch.tensor(table.copy()))
    opt = torch.optim.Adagrad([p], lr=lr, initial_accumulator_value=float(O.ADAGRAD_INIT_ACC), eps=float(O.KERAS_EPS))
    for _ in range(4):
        idx = rng.integers(0, rows, 14)                         # 14 draws from 9 rows: duplicates every step
        val = rng.standard_normal((14, e)).astype(np.float32)
        O.adagrad_sparse(table, acc, O.IndexedSlices(idx, val), lr)
        g = np.zeros((rows, e), np.float32)
        for i, v in zip(idx, val):                              # position order, fp32: the order dedup_indexed_slices defines
            g[i] += v
        O.adagrad_dense(dense_w, dense_acc, g, lr)
        p.grad = torch.tensor(g)
        opt.step()
        np.testing.assert_array_equal(table, dense_w)
        np.testing.assert_array_equal(acc, dense_acc)
        np.testing.assert_allclose(table, p.detach().numpy(), rtol=0, atol=1e-7)


def _trunc_f32(v):
    """float64 -> fp32 rounded TOWARD ZERO (the most pessimistic model of a tensor-core accumulator)."""
    f = v.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(v)
    f[over] = np.nextafter(f[over], np.float32(0))
    return f


@pytest.mark.parametrize("E", [32, 64, 128])
def test_index_filter_error_bound_covers_tf32_scores(E):
    """The exactness argument of the tensor-core index rests on |a_ij - s_ij| <= 2^-9 ||q_i|| ||c_j||, where a is the
    score of the TF32-rounded operands however the tensor core accumulates it and s the canonical fp32 score the oracle (and the
    exact rescoring kernel) computes.  Restated on the CPU: random rows, ReLU-like rows, and the adversarial case (parallel
    vectors whose every entry sits just below a TF32 rounding tie, so all 2E operand roundings push the score the same way)."""
    rng = np.random.default_rng(E)
    nq, n = 48, 96
    worst = np.float32(1.0 + (2 ** 12 - 1) * 2.0 ** -23)            # rounds up by (almost) half a TF32 ulp
    cases = {
        "gaussian": (rng.standard_normal((nq, E)), rng.standard_normal((n, E))),
        "relu": (np.abs(rng.standard_normal((nq, E))) * 0.1, np.maximum(rng.standard_normal((n, E)), 0) * 0.1),
        "adversarial": (np.full((nq, E), worst) * 2.0 ** rng.integers(-3, 3, (nq, 1)), np.full((n, E), worst) * 2.0 ** rng.integers(-3, 3, (n, 1))),
    }
    for name, (q, c) in cases.items():
        q, c = q.astype(np.float32), c.astype(np.float32)
        s = O.logits_qct(q, c, canonical=True).astype(np.float64)
        qt, ct = O.round_tf32(q).astype(np.float64), O.round_tf32(c).astype(np.float64)
        assert np.all(np.abs(qt - q) <= 2.0 ** -11 * np.abs(q) * (1 + 1e-6))
        bound = 2.0 ** -9 * np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(c.astype(np.float64), axis=1)[None, :]
        exact = qt @ ct.T                                               # products of TF32 values are exact in float64
        acc = np.zeros((nq, n), np.float32)                             # sequential accumulation, every add truncated
        for k in range(E):
            acc = _trunc_f32(acc.astype(np.float64) + qt[:, k:k + 1] * ct[None, :, k])
        for a in (exact, acc.astype(np.float64)):
            ratio = np.max(np.abs(a - s) / np.maximum(bound, 1e-300))
            assert ratio <= 0.55, (name, ratio)                         # 2^-10 from the operand roundings + accumulation, of 2^-9
        if name == "adversarial":
            assert np.max(np.abs(exact - s) / bound) > 0.45             # the case really is near the analytical worst (half the bound)


def test_tf32_rounded_values_convert_to_fp16_exactly_in_range():
    """The softmax kernels feed fp16 operand tiles made from TF32-rounded tower outputs: TF32 and fp16 carry the
    same 11-bit significand, so inside fp16's normal range the conversion is exact and the logits are those of the TF32 path.
    Below 2^-14 the value lands on fp16's subnormal grid (absolute error <= 2^-25), above 65504 it overflows -- stated limits."""
    rng = np.random.default_rng(3)
    mag = np.exp2(rng.uniform(-14, np.log2(65504.0), 200000)).astype(np.float32)
    x = O.round_tf32(mag * rng.choice([-1.0, 1.0], mag.shape).astype(np.float32))
    x = x[(np.abs(x) >= 2.0 ** -14) & (np.abs(x) <= 65504)]
    assert np.array_equal(x.astype(np.float16).astype(np.float32), x)
    tiny = O.round_tf32(np.exp2(rng.uniform(-30, -14, 50000)).astype(np.float32))
    assert np.max(np.abs(tiny.astype(np.float16).astype(np.float64) - tiny)) <= 2.0 ** -25
    with np.errstate(over="ignore"):
        assert np.isinf(np.float32(70000.0).astype(np.float16))


@pytest.mark.parametrize("K,group", [(1, 32), (12, 32), (100, 32), (12, 128)])
def test_index_filter_keeps_every_true_top_k_row(K, group):
    """The filter -> exact-rescore design of the tensor-core index, restated with numpy: per group of corpus rows
    keep the lower bound max_j a_ij - kappa_i max_j ||c_j||; lambda_i = any lower bound of the K-th largest group value; list the
    columns with a_ij + kappa_i ||c_j|| >= lambda_i.  Every row of the oracle's exact top-K (ties included) must be listed, for
    frequency-ordered corpora (best rows adjacent, hence the permutation), exact duplicates and signed rows alike."""
    rng = np.random.default_rng(100 * K + group)
    n, E, nq = 6400, 32, 24
    scale = 1.0 / (1.0 + np.arange(n) / 200.0)                         # popular rows first: large norms next to each other
    corpora = {
        "frequency-ordered": np.abs(rng.standard_normal((n, E))) * scale[:, None],
        "duplicates": np.repeat(np.abs(rng.standard_normal((n // 8, E))), 8, axis=0),
        "signed": rng.standard_normal((n, E)),
    }
    a_mult = int(0.618 * n) | 1
    while np.gcd(a_mult, n) != 1:
        a_mult += 2
    pos = (np.arange(n, dtype=np.int64) * a_mult) % n                   # fixed affine permutation of the prepared copy
    for name, c in corpora.items():
        c = c.astype(np.float32)
        q = (np.abs(rng.standard_normal((nq, E))) if name != "signed" else rng.standard_normal((nq, E))).astype(np.float32)
        s = O.logits_qct(q, c, canonical=True)
        _, top = O.top_k(s, K)
        a = (O.round_tf32(q).astype(np.float64) @ O.round_tf32(c).astype(np.float64).T).astype(np.float32)
        kappa = (2.0 ** -9 * np.linalg.norm(q.astype(np.float64), axis=1)).astype(np.float32)
        cn = np.linalg.norm(c.astype(np.float64), axis=1).astype(np.float32)
        counts = {}
        for layout, order in (("permuted", np.argsort(pos)), ("as stored", np.arange(n))):
            ap, cnp = a[:, order], cn[order]                            # corpus rows in the order the groups are cut from
            gv = ap.reshape(nq, n // group, group).max(axis=2) - kappa[:, None] * cnp.reshape(n // group, group).max(axis=1)[None, :]
            lam = np.sort(gv, axis=1)[:, -K]                            # K-th largest group value ...
            lam = lam - np.abs(lam) * 2.0 ** -12                        # ... or anything below it (the radix select stops at 16 bits)
            listed = a + kappa[:, None] * cn[None, :] >= lam[:, None]
            for i in range(nq):
                assert listed[i, top[i]].all(), (name, layout, i)       # exactness never depends on the layout
            counts[layout] = listed.sum() / nq
        dup = 8 if name == "duplicates" else 1
        assert counts["permuted"] <= 2 * K * dup + 16, (name, counts)   # ~K candidates reach the exact rescoring
        if name == "frequency-ordered" and K >= 12:                     # why the prepared copy is permuted (features.py:119-127)
            assert counts["as stored"] > 5 * counts["permuted"], counts
