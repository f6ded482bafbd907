"""tcgen05 / TMA / TMEM kernels against the oracle.  Operands are TF32-rounded first (tt_round_tf32), so the
tensor-core products are exact and only the fp32 accumulation order differs from the float64 oracle:
logits agree to ~1e-6 relative; the north-star bound (1e-3 relative for loss and logits) is asserted with
the tolerance written next to each check."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import two_tower_oracle as O  # noqa: E402

TC = 2


@pytest.fixture(scope="module")
def T(lib):
    import torch

    if not lib.tt_tc_available(0, 64):
        pytest.skip("tensor-core path unavailable on this device")
    return torch


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a)).cuda()


def stream():
    from pkg import _native as N

    return N.stream_ptr()


def round_tf32(lib, T, a):
    from pkg import _native as N

    src = dev(T, a)
    dst = T.empty_like(src)
    N.check(lib.tt_round_tf32(src.data_ptr(), a.shape[1], dst.data_ptr(), a.shape[1], a.shape[0], a.shape[1], stream()))
    return dst


@pytest.mark.parametrize("Bq,Bc,E", [(128, 128, 64), (1, 1, 32), (300, 200, 32), (257, 1000, 128), (1024, 2100, 64), (130, 64, 64)])
def test_logits_tc_matches_float64(lib, T, Bq, Bc, E):
    from pkg import _native as N

    rng = np.random.default_rng(1)
    q = rng.standard_normal((Bq, E)).astype(np.float32)
    c = rng.standard_normal((Bc, E)).astype(np.float32)
    bias = rng.standard_normal(Bc).astype(np.float32)
    dq, dc, db = round_tf32(lib, T, q), round_tf32(lib, T, c), dev(T, bias)
    qr, cr = dq.cpu().numpy(), dc.cpu().numpy()
    assert np.all((qr.view(np.uint32) & 0x1FFF) == 0)
    want = O.logits_qct(qr, cr).astype(np.float64) - bias[None, :]
    z = T.full((Bq, Bc), 777.0, dtype=T.float32, device="cuda")
    N.check(lib.tt_logits(dq.data_ptr(), E, dc.data_ptr(), E, db.data_ptr(), Bq, Bc, E, z.data_ptr(), Bc, TC, stream()))
    got = z.cpu().numpy()
    # exact TF32 products, fp32 accumulation of <= 128 terms: 1e-5 of the row scale
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * (np.abs(want).max() + 1.0))


def test_logits_tc_dyadic_bit_exact(lib, T):
    from pkg import _native as N

    rng = np.random.default_rng(2)
    q = (rng.integers(-16, 17, size=(200, 64)) / 16.0).astype(np.float32)      # exactly representable in TF32
    c = (rng.integers(-16, 17, size=(333, 64)) / 16.0).astype(np.float32)
    dq, dc = dev(T, q), dev(T, c)
    z = T.empty((200, 333), dtype=T.float32, device="cuda")
    N.check(lib.tt_logits(dq.data_ptr(), 64, dc.data_ptr(), 64, None, 200, 333, 64, z.data_ptr(), 333, TC, stream()))
    assert np.array_equal(z.cpu().numpy(), O.logits_qct(q, c, canonical=True))  # every path is exact on this grid


@pytest.mark.parametrize("Bq,Bc,E,off,bias", [(128, 128, 64, 0, True), (1000, 1000, 64, 0, True), (257, 257, 32, 0, True),
                                              (130, 390, 64, 130, True), (300, 300, 128, 0, False), (2048, 2048, 64, 0, True),
                                              (3, 3, 32, 0, True)])
def test_softmax_fwd_bwd_tc(lib, T, Bq, Bc, E, off, bias):
    from pkg import _native as N

    rng = np.random.default_rng(5)
    q = np.maximum(rng.standard_normal((Bq, E)) * 0.3, 0).astype(np.float32)
    c = np.maximum(rng.standard_normal((Bc, E)) * 0.3, 0).astype(np.float32)
    p = (rng.random(Bc) * 0.01 + 1e-5).astype(np.float32) if bias else None
    dq_, dc_ = round_tf32(lib, T, q), round_tf32(lib, T, c)
    qr, cr = dq_.cpu().numpy(), dc_.cpu().numpy()
    s = O.logits_qct(qr, cr)
    z = O.logq_correction(s, p) if bias else s
    loss, lse, dz = O.ce_sum_from_logits(z, diag_offset=off)
    want_dq, want_dc = dz @ cr.astype(np.float64), dz.T @ qr.astype(np.float64)
    dbias = T.log(dev(T, p)) if bias else None
    bp = dbias.data_ptr() if bias else None
    d_lse = T.empty(Bq, dtype=T.float32, device="cuda"); d_loss = T.zeros(1, dtype=T.float32, device="cuda")
    ws = T.empty(int(lib.tt_softmax_workspace_bytes(Bq, Bc, E)), dtype=T.uint8, device="cuda")
    N.check(lib.tt_inbatch_softmax_fwd(dq_.data_ptr(), E, dc_.data_ptr(), E, bp, Bq, Bc, E, off, d_lse.data_ptr(), d_loss.data_ptr(),
                                       ws.data_ptr(), ws.numel(), TC, stream()))
    # north star: loss within 1e-3 relative; measured agreement is ~1e-6, assert 1e-5
    assert abs(float(d_loss) - loss) <= 1e-5 * abs(loss) + 1e-5
    np.testing.assert_allclose(d_lse.cpu().numpy(), lse, rtol=1e-5, atol=1e-5)
    gq = T.full((Bq, E), 9.0, dtype=T.float32, device="cuda"); gc = T.full((Bc, E), 9.0, dtype=T.float32, device="cuda")
    N.check(lib.tt_inbatch_softmax_bwd(dq_.data_ptr(), E, dc_.data_ptr(), E, bp, d_lse.data_ptr(), Bq, Bc, E, off, gq.data_ptr(), E,
                                       gc.data_ptr(), E, ws.data_ptr(), ws.numel(), TC, stream()))
    # dZ is rounded to TF32 (2^-11 relative per element) before the second MMA: 1e-3 of the gradient scale
    np.testing.assert_allclose(gq.cpu().numpy(), want_dq, rtol=0, atol=1e-3 * (np.abs(want_dq).max() + 1e-6))
    np.testing.assert_allclose(gc.cpu().numpy(), want_dc, rtol=0, atol=1e-3 * (np.abs(want_dc).max() + 1e-6))
    # deterministic: fixed-order split reduction, no atomics
    gq2 = T.empty_like(gq); gc2 = T.empty_like(gc)
    N.check(lib.tt_inbatch_softmax_bwd(dq_.data_ptr(), E, dc_.data_ptr(), E, bp, d_lse.data_ptr(), Bq, Bc, E, off, gq2.data_ptr(), E,
                                       gc2.data_ptr(), E, ws.data_ptr(), ws.numel(), TC, stream()))
    assert T.equal(gq, gq2) and T.equal(gc, gc2)


def test_tc_and_exact_paths_agree_on_raw_fp32_operands(lib, T):
    """Un-rounded fp32 operands: the tensor core truncates/rounds them to TF32 itself; the north-star bound
    (logits within 1e-3 relative of the fp32 reference) must still hold."""
    from pkg import _native as N

    rng = np.random.default_rng(6)
    Bq = Bc = 512; E = 64
    q = np.maximum(rng.standard_normal((Bq, E)) * 0.3, 0).astype(np.float32)
    c = np.maximum(rng.standard_normal((Bc, E)) * 0.3, 0).astype(np.float32)
    dq_, dc_ = dev(T, q), dev(T, c)
    z_tc = T.empty((Bq, Bc), dtype=T.float32, device="cuda"); z_ex = T.empty_like(z_tc)
    N.check(lib.tt_logits(dq_.data_ptr(), E, dc_.data_ptr(), E, None, Bq, Bc, E, z_tc.data_ptr(), Bc, TC, stream()))
    N.check(lib.tt_logits(dq_.data_ptr(), E, dc_.data_ptr(), E, None, Bq, Bc, E, z_ex.data_ptr(), Bc, 1, stream()))
    a, b = z_tc.cpu().numpy(), z_ex.cpu().numpy()
    rel = np.abs(a - b).max() / np.abs(b).max()
    assert rel < 1e-3, rel


def test_model_train_step_tc_within_north_star(lib, T):
    from pkg import _native as N
    from pkg.modelling._device import set_seed
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory
    from pkg.schema import dtypes as tt
    from pkg.schema.features import Feature, FeatureFamily

    def build(impl):
        set_seed(3)
        qf = [Feature("age", tt.float32, FeatureFamily.QUERY), Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=64)]
        cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=64)]
        qf[1].set_vocab_size(4000); cf[0].set_vocab_size(900)
        m = TwoTowerModel(qf, cf, "article_id", 64, candidate_prob_lookup={str(i + 1): 1.0 / 900 for i in range(900)})
        m.impl = impl
        m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
        return m

    rng = np.random.default_rng(9)
    B = 1500
    batches = [{"age": rng.random((B, 1)).astype(np.float32), "customer_id": rng.integers(0, 4001, size=(B, 1)).astype(np.int32),
                "article_id": np.minimum(rng.zipf(1.2, size=(B, 1)), 900).astype(np.int32)} for _ in range(4)]
    exact, tc = build(N.TT_IMPL_SIMT), build(N.TT_IMPL_AUTO)
    assert tc._tc_ok()
    for b in batches:
        le, lt = float(exact.train_step(b)["loss"]), float(tc.train_step(b)["loss"])
        assert abs(le - lt) <= 1e-3 * abs(le), (le, lt)                   # north star: loss within 1e-3 relative
    we = exact.candidate_tower.input_layer.embedding_layers["article_id"].weight
    wt = tc.candidate_tower.input_layer.embedding_layers["article_id"].weight
    # trajectories stay together over 4 steps (Adagrad's early steps are ~lr*sign(g), so isolated elements whose
    # gradient is ~0 may differ by a few 1e-2; the bulk must agree)
    assert float((we - wt).abs().mean()) < 5e-4
    assert float(((we - wt).abs() > 1e-2).float().mean()) < 1e-3


# ---------------------------------------------------------------------------------------------------
# index: tensor-core filter + exact rescoring must return exactly what the canonical fp32 oracle returns
# ---------------------------------------------------------------------------------------------------
def _index_tc(lib, T, q, c, K, idx_base=0, prepared=False):
    from pkg import _native as N

    dq, dc = dev(T, q), dev(T, c)
    nq, n, E = q.shape[0], c.shape[0], q.shape[1]
    c32 = mx = None
    if prepared:   # what BruteForceIndex does once at build time
        rows_pad = ((n + 255) // 256 + 1) * 256
        n_pad = 2 * rows_pad + rows_pad // 32 + 32                 # TT_INDEX_NORM_PAD
        c32 = T.empty_like(dc); mx = T.full((n_pad,), 9.0, dtype=T.float32, device="cuda")
        N.check(lib.tt_index_prepare(dc.data_ptr(), E, n, E, c32.data_ptr(), mx.data_ptr(), stream()))
    s = T.empty((nq, K), dtype=T.float32, device="cuda"); i = T.empty((nq, K), dtype=T.int32, device="cuda")
    ws = T.empty(int(lib.tt_index_workspace_bytes(nq, n, E, K, TC, 1 if prepared else 0)), dtype=T.uint8, device="cuda")
    N.check(lib.tt_index_topk(dq.data_ptr(), E, dc.data_ptr(), E, c32.data_ptr() if prepared else None, mx.data_ptr() if prepared else None,
                              nq, n, E, K, idx_base, s.data_ptr(), i.data_ptr(), ws.data_ptr(), ws.numel(), TC, stream()))
    return s.cpu().numpy(), i.cpu().numpy()


@pytest.mark.parametrize("nq,n,E,K,prepared", [(70, 20_000, 64, 100, False), (2048, 105_542, 64, 100, True), (5, 8000, 32, 12, True),
                                               (300, 40_000, 128, 100, False), (129, 50_001, 64, 1, True), (64, 140_000, 64, 500, True)])
def test_index_tc_bit_exact(lib, T, nq, n, E, K, prepared):
    rng = np.random.default_rng(10)
    q = np.maximum(rng.standard_normal((nq, E)) * 0.3, 0).astype(np.float32)      # tower outputs are non-negative
    c = (np.abs(rng.standard_normal((n, E))) * 0.1).astype(np.float32)
    want_s, want_i = O.index_topk(q, c, K, idx_base=7)
    s, i = _index_tc(lib, T, q, c, K, idx_base=7, prepared=prepared)
    assert np.array_equal(i, want_i.astype(np.int32))                              # indices bit-exact
    assert np.array_equal(s, want_s)                                               # scores are the canonical fp32 values


def test_index_tc_signed_embeddings_and_many_exact_ties(lib, T):
    rng = np.random.default_rng(11)
    # signed values (BruteForceIndex accepts any query model) on a coarse dyadic grid: thousands of exactly equal
    # scores, so the (score desc, index asc) rule decides most of the top-100
    q = (rng.integers(-4, 5, size=(40, 32)) / 4.0).astype(np.float32)
    c = (rng.integers(-2, 3, size=(30_000, 32)) / 2.0).astype(np.float32)
    want_s, want_i = O.index_topk(q, c, 100)
    s, i = _index_tc(lib, T, q, c, 100)
    assert np.array_equal(i, want_i.astype(np.int32)) and np.array_equal(s, want_s)


def test_index_tc_overflow_falls_back_to_exact_on_device(lib, T):
    rng = np.random.default_rng(12)
    q = np.maximum(rng.standard_normal((200, 64)) * 0.3, 0).astype(np.float32)
    c = (np.abs(rng.standard_normal((20_000, 64))) * 0.1).astype(np.float32)
    want_s, want_i = O.index_topk(q, c, 100)
    lib.tt_debug_index_cap(64)          # candidate lists of 64 < K: every query overflows
    try:
        s, i = _index_tc(lib, T, q, c, 100)
    finally:
        lib.tt_debug_index_cap(0)
    assert np.array_equal(i, want_i.astype(np.int32)) and np.array_equal(s, want_s)


def test_brute_force_index_uses_tc_and_matches_exact(lib, T):
    from pkg import _native as N
    from pkg.modelling.indices.brute_force import BruteForceIndex

    rng = np.random.default_rng(13)
    n, E = 50_000, 64
    c = (np.abs(rng.standard_normal((n, E))) * 0.1).astype(np.float32)
    table = np.maximum(rng.standard_normal((500, E)) * 0.3, 0).astype(np.float32)

    class QM:
        def __call__(self, x):
            return table[np.asarray(x["id"]).reshape(-1)]

        def get_input_signature(self):
            return {}

    ids = np.arange(n, dtype=np.int32)
    index = BruteForceIndex(100, QM(), [(ids[:30_000], c[:30_000]), (ids[30_000:], c[30_000:])])
    assert index._candidates_tf32 is not None
    x = {"id": rng.integers(0, 500, size=(333, 1)).astype(np.int32)}
    got = index(x)
    index.impl = N.TT_IMPL_SIMT
    assert np.array_equal(got, index(x))
    _, want = O.index_topk(table[x["id"].reshape(-1)], c, 100)
    assert np.array_equal(got, ids[want])


@pytest.mark.parametrize("Bq,Bc,E,off", [(1000, 1000, 64, 0), (257, 300, 32, 5), (130, 777, 128, 3)])
def test_softmax_step_equals_fwd_then_bwd(lib, T, Bq, Bc, E, off):
    """tt_inbatch_softmax_step against the two separate calls.  E = 32 (stream-K kernels, shared prep): bit-identical lse / dQ / dC.
    E >= 64 (two-pass kernels): the step forms dQ inside the forward sweep from the un-normalised weights, the backward entry point
    from the normalised ones -- same lse bit for bit, gradients equal to the fp16 rounding of the weights (2^-11 of the row scale)."""
    from pkg import _native as N

    rng = np.random.default_rng(11)
    q = np.maximum(rng.standard_normal((Bq, E)) * 0.3, 0).astype(np.float32)
    c = np.maximum(rng.standard_normal((Bc, E)) * 0.3, 0).astype(np.float32)
    p = (rng.random(Bc) * 0.01 + 1e-5).astype(np.float32)
    dq_, dc_ = T.from_numpy(q).cuda(), T.from_numpy(c).cuda()
    q32, c32 = T.empty_like(dq_), T.empty_like(dc_)
    N.check(lib.tt_round_tf32(dq_.data_ptr(), E, q32.data_ptr(), E, Bq, E, stream()))
    N.check(lib.tt_round_tf32(dc_.data_ptr(), E, c32.data_ptr(), E, Bc, E, stream()))
    bias = T.log(T.from_numpy(p).cuda())
    ws = T.empty(int(lib.tt_softmax_workspace_bytes(Bq, Bc, E)), dtype=T.uint8, device="cuda")
    out = {}
    for name in ("sep", "step"):
        lse = T.zeros(Bq, device="cuda"); loss = T.zeros(1, device="cuda")
        gq = T.zeros((Bq, E), device="cuda"); gc = T.zeros((Bc, E), device="cuda")
        if name == "sep":
            N.check(lib.tt_inbatch_softmax_fwd(q32.data_ptr(), E, c32.data_ptr(), E, bias.data_ptr(), Bq, Bc, E, off, lse.data_ptr(), loss.data_ptr(),
                                               ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, stream()))
            N.check(lib.tt_inbatch_softmax_bwd(q32.data_ptr(), E, c32.data_ptr(), E, bias.data_ptr(), lse.data_ptr(), Bq, Bc, E, off, gq.data_ptr(), E,
                                               gc.data_ptr(), E, ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, stream()))
        else:
            N.check(lib.tt_inbatch_softmax_step(q32.data_ptr(), E, c32.data_ptr(), E, bias.data_ptr(), Bq, Bc, E, off, lse.data_ptr(), loss.data_ptr(),
                                                gq.data_ptr(), E, gc.data_ptr(), E, ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, stream()))
        out[name] = [x.cpu().numpy() for x in (lse, loss, gq, gc)]
    assert np.array_equal(out["sep"][0], out["step"][0])
    assert abs(out["sep"][1][0] - out["step"][1][0]) <= 1e-6 * abs(out["sep"][1][0])
    for a, b in zip(out["sep"][2:], out["step"][2:]):
        if E == 32:
            assert np.array_equal(a, b)
        else:
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-3 * np.abs(a).max())


def _softmax_step(lib, T, q, c, bias, off=0):
    from pkg import _native as N

    Bq, E = q.shape
    Bc = c.shape[0]
    dq_, dc_ = dev(T, q), dev(T, c)
    db = dev(T, bias) if bias is not None else None
    lse = T.zeros(Bq, device="cuda"); loss = T.zeros(1, device="cuda")
    gq = T.full((Bq, E), 9.0, device="cuda"); gc = T.full((Bc, E), 9.0, device="cuda")
    ws = T.empty(int(lib.tt_softmax_workspace_bytes(Bq, Bc, E)), dtype=T.uint8, device="cuda")
    N.check(lib.tt_inbatch_softmax_step(dq_.data_ptr(), E, dc_.data_ptr(), E, db.data_ptr() if db is not None else None, Bq, Bc, E, off, lse.data_ptr(),
                                        loss.data_ptr(), gq.data_ptr(), E, gc.data_ptr(), E, ws.data_ptr(), ws.numel(), TC, stream()))
    return float(loss.cpu()[0]), lse.cpu().numpy(), gq.cpu().numpy(), gc.cpu().numpy()


def _row_err(got, want):
    """per-row error relative to the row norm; rows whose norm is below 1e-3 of the largest are measured against that floor"""
    n = np.linalg.norm(want, axis=1)
    return float((np.linalg.norm(got - want, axis=1) / np.maximum(n, 1e-3 * n.max())).max())


def test_softmax_step_c2_shape_vs_fp32_oracle_on_unrounded_inputs(lib, T):
    """The bench's configuration (B = 8192, E = 64, logQ correction) on the tensor-core path against the oracle evaluated on the SAME
    un-rounded fp32 operands (float64 arithmetic): north-star tolerance -- loss within 1e-3 relative (measured ~1e-6), lse within 1e-3,
    every row of dQ and dC within 1e-3 of its norm."""
    rng = np.random.default_rng(21)
    B, E = 8192, 64
    q = np.maximum(rng.standard_normal((B, E)) * 0.3, 0).astype(np.float32)
    c = np.maximum(rng.standard_normal((B, E)) * 0.3, 0).astype(np.float32)
    p = (rng.random(B) * 0.01 + 1e-5).astype(np.float32)
    z = O.logits_qct(q, c).astype(np.float64) - np.log(p.astype(np.float64))[None, :]      # float64 product of the fp32 operands
    loss, lse, dz = O.ce_sum_from_logits(z)
    want_dq, want_dc = dz @ c.astype(np.float64), dz.T @ q.astype(np.float64)
    got_loss, got_lse, gq, gc = _softmax_step(lib, T, q, c, np.log(p))
    assert abs(got_loss - loss) <= 1e-3 * abs(loss)
    assert abs(got_loss - loss) <= 2e-5 * abs(loss)            # (what the fp16/TF32-class operands actually deliver)
    np.testing.assert_allclose(got_lse, lse, rtol=0, atol=1e-3)
    assert _row_err(gq, want_dq) <= 1e-3
    assert _row_err(gc, want_dc) <= 1e-3


@pytest.mark.parametrize("qs,cs", [(3.0e5, 1.0e-5), (2.0e-6, 2.0e5), (1.0, 1.0)])
def test_softmax_step_operands_outside_the_fp16_range(lib, T, qs, cs):
    """Tower outputs above 65504 or below 2^-14 (un-normalised numeric features, a diverging run): the operands are scaled per tensor by
    a power of two before the fp16 conversion, so nothing saturates or flushes -- results stay within the north-star tolerance of the
    float64 oracle, and nothing is NaN / inf."""
    rng = np.random.default_rng(22)
    Bq, Bc, E = 700, 900, 64
    q = (np.maximum(rng.standard_normal((Bq, E)) * 0.3, 0) * qs).astype(np.float32)
    c = (np.maximum(rng.standard_normal((Bc, E)) * 0.3, 0) * cs).astype(np.float32)
    assert qs == 1.0 or q.max() > 65504 or c.max() > 65504
    p = (rng.random(Bc) * 0.01 + 1e-5).astype(np.float32)
    z = O.logits_qct(q, c).astype(np.float64) - np.log(p.astype(np.float64))[None, :]
    loss, lse, dz = O.ce_sum_from_logits(z, diag_offset=100)
    want_dq, want_dc = dz @ c.astype(np.float64), dz.T @ q.astype(np.float64)
    got_loss, got_lse, gq, gc = _softmax_step(lib, T, q, c, np.log(p), off=100)
    assert np.isfinite(got_loss) and np.isfinite(gq).all() and np.isfinite(gc).all()
    assert abs(got_loss - loss) <= 1e-3 * abs(loss)
    np.testing.assert_allclose(got_lse, lse, rtol=0, atol=1e-3 * max(1.0, np.abs(lse).max()))
    assert _row_err(gq, want_dq) <= 2e-3
    assert _row_err(gc, want_dc) <= 2e-3


def test_softmax_step_sharp_softmax_keeps_the_positive_term_exact(lib, T):
    """A trained model: the positive dominates its row, so dQ_i = sum_j p_ij c_j - c_i is a difference of nearly equal vectors.  The
    positive is kept out of the tensor-core products and its term (p_ii - 1) is formed without cancellation in fp32: every row stays
    within 1e-2 of its norm although the logits themselves carry the operand rounding (~1e-4 relative)."""
    rng = np.random.default_rng(23)
    B, E = 1000, 64
    q = np.maximum(rng.standard_normal((B, E)) * 0.3, 0).astype(np.float32)
    c = (q + 0.05 * np.maximum(rng.standard_normal((B, E)) * 0.3, 0)).astype(np.float32)
    q = (q * 6.0).astype(np.float32)
    p = (rng.random(B) * 0.01 + 1e-5).astype(np.float32)
    z = O.logits_qct(q, c).astype(np.float64) - np.log(p.astype(np.float64))[None, :]
    loss, lse, dz = O.ce_sum_from_logits(z)
    want_dq, want_dc = dz @ c.astype(np.float64), dz.T @ q.astype(np.float64)
    got_loss, got_lse, gq, gc = _softmax_step(lib, T, q, c, np.log(p))
    assert abs(got_loss - loss) <= 1e-3 * abs(loss)
    assert _row_err(gq, want_dq) <= 1e-2
    assert _row_err(gc, want_dc) <= 1e-2
