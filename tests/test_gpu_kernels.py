"""Parity of every C-ABI entry point against the oracle, on the GPU.  Integer / index work must be
bit-exact; fp32 contractions on the exact (SIMT) path are bit-exact against the canonical-order C oracle;
tolerances for the remaining floating-point pieces are written next to each assertion."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import two_tower_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def T():
    import torch

    return torch


def dev(T, a, dtype=None):
    t = T.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def feats_array(N, entries):
    from pkg.modelling._device import feature_array

    return feature_array(entries)


def stream():
    from pkg import _native as N

    return N.stream_ptr()


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec", [
    dict(B=257, num=0, embs=[32]),                 # C1 tower: id only, vector path
    dict(B=1000, num=0, embs=[64, 16, 8]),         # C2 candidate tower, vector path
    dict(B=333, num=1, embs=[64]),                 # C2 query tower: numeric first -> scalar path
    dict(B=5, num=2, embs=[3, 2]),                 # odd widths (main.py uses e=2)
    dict(B=0, num=0, embs=[4]),                    # empty batch
])
def test_gather_concat_bit_exact(lib, T, spec):
    from pkg import _native as N

    rng = np.random.default_rng(1)
    B = spec["B"]
    tables = [rng.standard_normal((50 + 7 * i, e)).astype(np.float32) for i, e in enumerate(spec["embs"])]
    ids = [rng.integers(0, t.shape[0], size=B).astype(np.int32) for t in tables]
    if B:
        ids[0][0] = 0                                   # OOV row
    nums = [rng.random(B).astype(np.float32) for _ in range(spec["num"])]
    want = np.concatenate([n.reshape(-1, 1) for n in nums] + [t[i] for t, i in zip(tables, ids)], axis=1) if B else None
    D = spec["num"] + sum(spec["embs"])
    ld = (D + 3) // 4 * 4
    d_tab = [dev(T, t) for t in tables]
    d_ids = [dev(T, i) for i in ids]
    d_num = [dev(T, n) for n in nums]
    entries, col = [], 0
    for n in d_num:
        entries.append(dict(table=None, src=n.data_ptr(), rows=0, e=1, col=col)); col += 1
    for t, i in zip(d_tab, d_ids):
        entries.append(dict(table=t.data_ptr(), src=i.data_ptr(), rows=t.shape[0], e=t.shape[1], col=col)); col += t.shape[1]
    X = T.full((max(B, 1), ld), -7.0, dtype=T.float32, device="cuda")
    N.check(lib.tt_gather_concat(feats_array(N, entries), len(entries), B, D, X.data_ptr(), ld, stream()))
    T.cuda.synchronize()
    if B:
        got = X.cpu().numpy()
        assert np.array_equal(got[:, :D], want)
        assert np.all(got[:, D:] == 0)


def test_gather_out_of_range_id_is_oov(lib, T):
    from pkg import _native as N

    t = dev(T, np.arange(12, dtype=np.float32).reshape(3, 4))
    ids = dev(T, np.array([2, 3, -1, 99], np.int32))
    X = T.zeros((4, 4), dtype=T.float32, device="cuda")
    fa = feats_array(N, [dict(table=t.data_ptr(), src=ids.data_ptr(), rows=3, e=4, col=0)])
    N.check(lib.tt_gather_concat(fa, 1, 4, 4, X.data_ptr(), 4, stream()))
    assert X.cpu().numpy().tolist() == [[8, 9, 10, 11], [0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3]]


@pytest.mark.parametrize("B,K,N_", [(300, 65, 64), (64, 88, 64), (1, 32, 32), (513, 258, 256), (70, 7, 5)])
def test_dense_fwd_bit_exact_vs_canonical_oracle(lib, T, B, K, N_):
    from pkg import _native as N

    rng = np.random.default_rng(2)
    x = rng.standard_normal((B, K)).astype(np.float32)
    w = (rng.standard_normal((K, N_)) * 0.2).astype(np.float32)
    b = rng.standard_normal(N_).astype(np.float32)
    want = O.dense_relu(x, w, b, canonical=True)
    dx, dw, db = dev(T, x), dev(T, w), dev(T, b)
    y = T.empty((B, N_), dtype=T.float32, device="cuda")
    y32 = T.empty((B, N_), dtype=T.float32, device="cuda")
    N.check(lib.tt_dense_fwd(dx.data_ptr(), K, dw.data_ptr(), db.data_ptr(), y.data_ptr(), N_, y32.data_ptr(), B, K, N_, 1, stream()))
    got = y.cpu().numpy()
    assert np.array_equal(got, want)
    # TF32 copy: 10 explicit mantissa bits, round to nearest => |rel err| <= 2^-11, low 13 bits clear
    g32 = y32.cpu().numpy()
    assert np.all((g32.view(np.uint32) & 0x1FFF) == 0)
    np.testing.assert_allclose(g32, got, rtol=2.0 ** -11, atol=0)


def test_input_dense_fused_equals_gather_then_dense(lib, T):
    from pkg import _native as N

    rng = np.random.default_rng(3)
    B, embs, n_out = 515, [64, 16, 8], 64
    tables = [rng.standard_normal((40, e)).astype(np.float32) for e in embs]
    ids = [rng.integers(0, 40, size=B).astype(np.int32) for _ in embs]
    age = rng.random(B).astype(np.float32)
    D = 1 + sum(embs); ld = (D + 3) // 4 * 4
    w = (rng.standard_normal((D, n_out)) * 0.1).astype(np.float32); b = rng.standard_normal(n_out).astype(np.float32)
    x = np.concatenate([age.reshape(-1, 1)] + [t[i] for t, i in zip(tables, ids)], axis=1)
    want = O.dense_relu(x, w, b, canonical=True)
    keep = [dev(T, t) for t in tables] + [dev(T, i) for i in ids] + [dev(T, age)]
    entries, col = [dict(table=None, src=keep[-1].data_ptr(), rows=0, e=1, col=0)], 1
    for t, i in zip(keep[:3], keep[3:6]):
        entries.append(dict(table=t.data_ptr(), src=i.data_ptr(), rows=40, e=t.shape[1], col=col)); col += t.shape[1]
    X = T.full((B, ld), 3.0, dtype=T.float32, device="cuda")
    Y = T.empty((B, n_out), dtype=T.float32, device="cuda")
    dw_, db_ = dev(T, w), dev(T, b)
    N.check(lib.tt_input_dense_fwd(feats_array(N, entries), 4, D, dw_.data_ptr(), db_.data_ptr(), X.data_ptr(), ld,
                                   Y.data_ptr(), n_out, None, B, n_out, 1, stream()))
    assert np.array_equal(Y.cpu().numpy(), want)
    gx = X.cpu().numpy()
    assert np.array_equal(gx[:, :D], x) and np.all(gx[:, D:] == 0)


@pytest.mark.parametrize("B,K,N_", [(1000, 65, 64), (4096, 88, 64), (33, 258, 256), (1, 4, 4)])
def test_dense_bwd_vs_float64(lib, T, B, K, N_):
    from pkg import _native as N

    rng = np.random.default_rng(4)
    x = rng.standard_normal((B, K)).astype(np.float32)
    w = (rng.standard_normal((K, N_)) * 0.2).astype(np.float32)
    b = rng.standard_normal(N_).astype(np.float32)
    y = O.dense_relu(x, w, b, canonical=True)
    dy = rng.standard_normal((B, N_)).astype(np.float32)
    dpre = dy.astype(np.float64) * (y > 0)
    want_dw, want_db, want_dx = x.astype(np.float64).T @ dpre, dpre.sum(0), dpre @ w.astype(np.float64).T
    ws = T.empty(int(lib.tt_dense_bwd_workspace_bytes(B, K, N_)), dtype=T.uint8, device="cuda")
    dX = T.empty((B, K), dtype=T.float32, device="cuda"); dW = T.empty((K, N_), dtype=T.float32, device="cuda")
    dB = T.empty(N_, dtype=T.float32, device="cuda")
    args = (dev(T, x), dev(T, w), dev(T, y), dev(T, dy))
    def run():
        N.check(lib.tt_dense_bwd(args[0].data_ptr(), K, args[1].data_ptr(), args[2].data_ptr(), N_, args[3].data_ptr(), N_, dX.data_ptr(), K,
                                 dW.data_ptr(), dB.data_ptr(), B, K, N_, 1, ws.data_ptr(), ws.numel(), stream()))
        return dX.cpu().numpy().copy(), dW.cpu().numpy().copy(), dB.cpu().numpy().copy()
    gx, gw, gb = run()
    # fp32 accumulation over B (<= 4096) terms of O(1): relative 1e-5 of the column scale is ample
    scale_w = np.abs(want_dw).max() + 1e-6
    np.testing.assert_allclose(gw, want_dw, rtol=0, atol=2e-5 * scale_w)
    np.testing.assert_allclose(gb, want_db, rtol=0, atol=2e-5 * (np.abs(want_db).max() + 1e-6))
    np.testing.assert_allclose(gx, want_dx, rtol=0, atol=2e-5 * (np.abs(want_dx).max() + 1e-6))
    gx2, gw2, gb2 = run()                                   # deterministic: fixed-order batch reduction
    assert np.array_equal(gw, gw2) and np.array_equal(gb, gb2) and np.array_equal(gx, gx2)


# ---------------------------------------------------------------------------------------------------
def _softmax_case(rng, Bq, Bc, E, off, with_bias=True, scale=0.3):
    q = np.maximum(rng.standard_normal((Bq, E)) * scale, 0).astype(np.float32)   # towers end in ReLU: non-negative
    c = np.maximum(rng.standard_normal((Bc, E)) * scale, 0).astype(np.float32)
    p = rng.random(Bc).astype(np.float32) * 0.01 + 1e-5 if with_bias else None
    return q, c, p


IMPLS = [1]  # TT_IMPL_SIMT; the tcgen05 path is added by test_gpu_tc.py


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("Bq,Bc,E,off,bias", [(3, 3, 2, 0, True), (200, 200, 32, 0, True), (257, 257, 64, 0, True),
                                              (130, 390, 64, 130, True), (64, 64, 128, 0, False), (1, 1, 32, 0, True)])
def test_inbatch_softmax_fwd_bwd(lib, T, impl, Bq, Bc, E, off, bias):
    from pkg import _native as N

    rng = np.random.default_rng(5)
    q, c, p = _softmax_case(rng, Bq, Bc, E, off, bias)
    s = O.logits_qct(q, c)
    z = O.logq_correction(s, p) if bias else s
    loss, lse, dz = O.ce_sum_from_logits(z, diag_offset=off)
    want_dq, want_dc = dz @ c.astype(np.float64), dz.T @ q.astype(np.float64)
    dq_, dc_ = dev(T, q), dev(T, c)
    dbias = T.log(dev(T, p)) if bias else None
    d_lse = T.empty(Bq, dtype=T.float32, device="cuda"); d_loss = T.zeros(1, dtype=T.float32, device="cuda")
    ws = T.empty(int(lib.tt_softmax_workspace_bytes(Bq, Bc, E)), dtype=T.uint8, device="cuda")
    bp = dbias.data_ptr() if bias else None
    N.check(lib.tt_inbatch_softmax_fwd(dq_.data_ptr(), E, dc_.data_ptr(), E, bp, Bq, Bc, E, off, d_lse.data_ptr(), d_loss.data_ptr(),
                                       ws.data_ptr(), ws.numel(), impl, stream()))
    # north star: loss and logits within 1e-3 relative (fp32 path is far tighter: 1e-5)
    assert abs(float(d_loss) - loss) <= 1e-5 * abs(loss) + 1e-5
    np.testing.assert_allclose(d_lse.cpu().numpy(), lse, rtol=1e-5, atol=1e-5)
    gq = T.empty((Bq, E), dtype=T.float32, device="cuda"); gc = T.empty((Bc, E), dtype=T.float32, device="cuda")
    N.check(lib.tt_inbatch_softmax_bwd(dq_.data_ptr(), E, dc_.data_ptr(), E, bp, d_lse.data_ptr(), Bq, Bc, E, off, gq.data_ptr(), E,
                                       gc.data_ptr(), E, ws.data_ptr(), ws.numel(), impl, stream()))
    np.testing.assert_allclose(gq.cpu().numpy(), want_dq, rtol=0, atol=2e-5 * (np.abs(want_dq).max() + 1e-6))
    np.testing.assert_allclose(gc.cpu().numpy(), want_dc, rtol=0, atol=2e-5 * (np.abs(want_dc).max() + 1e-6))


def test_logits_and_logq_reference_fixture(lib, T, golden):
    from pkg import _native as N

    fix, _ = golden
    g = fix["logq"]
    logits = dev(T, np.array(g["logits"], np.float32))
    p = dev(T, np.array([g["candidate_prob_lookup"][i] for i in g["candidate_ids"]], np.float32))
    lp = T.empty_like(p); out = T.empty_like(logits)
    N.check(lib.tt_log_f32(p.data_ptr(), lp.data_ptr(), 3, stream()))
    N.check(lib.tt_logq_apply(logits.data_ptr(), 3, lp.data_ptr(), 3, 3, out.data_ptr(), 3, stream()))
    np.testing.assert_allclose(out.cpu().numpy(), np.array(g["expected"]), rtol=0, atol=5e-7)   # tests/test_layers.py:28-36
    # tt_logits: exact path is bit-identical to the canonical oracle
    rng = np.random.default_rng(6)
    q = rng.standard_normal((70, 48)).astype(np.float32); c = rng.standard_normal((90, 48)).astype(np.float32)
    z = T.empty((70, 90), dtype=T.float32, device="cuda")
    dq_, dc_ = dev(T, q), dev(T, c)
    N.check(lib.tt_logits(dq_.data_ptr(), 48, dc_.data_ptr(), 48, None, 70, 90, 48, z.data_ptr(), 90, 1, stream()))
    assert np.array_equal(z.cpu().numpy(), O.logits_qct(q, c, canonical=True))


# ---------------------------------------------------------------------------------------------------
def _sparse_jobs(N, T, specs, B, rng, adam=False):
    """specs: list of (rows, e, nsrc).  Returns ctypes jobs, device tensors to keep alive, and host copies."""
    jobs = (N.TTSparseJob * len(specs))()
    keep, host = [], []
    for j, (rows, e, nsrc) in enumerate(specs):
        table = rng.standard_normal((rows, e)).astype(np.float32)
        s0 = np.full((rows, e), 0.1, np.float32) if not adam else (rng.standard_normal((rows, e)) * 0.01).astype(np.float32)
        s1 = (rng.random((rows, e)) * 0.01).astype(np.float32)
        ld = e + 4
        idl, gl = [], []
        for s in range(nsrc):
            # zipf-like ids: heavy duplicates, plus an out-of-range id that must fold into row 0
            ids = np.minimum(rng.zipf(1.3, size=B) - 1, rows - 1).astype(np.int32)
            if B > 3:
                ids[3] = rows + 5
            g = rng.standard_normal((B, ld)).astype(np.float32)
            idl.append(ids); gl.append(g)
        dt, d0, d1 = dev(T, table), dev(T, s0), dev(T, s1)
        dids = [dev(T, i) for i in idl]; dg = [dev(T, g) for g in gl]
        keep += [dt, d0, d1] + dids + dg
        jobs[j].table, jobs[j].slot0, jobs[j].slot1 = dt.data_ptr(), d0.data_ptr(), d1.data_ptr()
        jobs[j].rows, jobs[j].e, jobs[j].nsrc, jobs[j].n_per_src = rows, e, nsrc, B
        for s in range(nsrc):
            jobs[j].ids[s] = dids[s].data_ptr(); jobs[j].grad[s] = dg[s].data_ptr() + 4 * 2; jobs[j].grad_ld[s] = ld
        ids_all = np.concatenate([np.where((i < 0) | (i >= rows), 0, i) for i in idl])
        vals_all = np.concatenate([g[:, 2:2 + e] for g in gl], axis=0)
        host.append((table, s0, s1, O.IndexedSlices(ids_all, vals_all), dt, d0, d1))
    return jobs, keep, host


@pytest.mark.parametrize("B", [1, 100, 2048, 5000])
def test_sparse_adagrad_bit_exact_and_deterministic(lib, T, B):
    from pkg import _native as N

    rng = np.random.default_rng(7)
    specs = [(1_000_003, 32, 1), (105_543, 64, 1), (132, 16, 2), (51, 8, 1), (70_000, 3, 1)]
    jobs, keep, host = _sparse_jobs(N, T, specs, B, rng)
    ws = T.empty(int(lib.tt_sparse_workspace_bytes(len(specs), 2 * B, max(s[1] for s in specs))), dtype=T.uint8, device="cuda")
    N.check(lib.tt_sparse_sort(jobs, len(specs), ws.data_ptr(), ws.numel(), stream()))
    N.check(lib.tt_sparse_adagrad(jobs, len(specs), 0.05, 1e-7, ws.data_ptr(), ws.numel(), stream()))
    T.cuda.synchronize()
    for table, acc, _, slices, dt, d0, _ in host:
        O.adagrad_sparse(table, acc, slices, 0.05)
        assert np.array_equal(dt.cpu().numpy(), table)          # bit-exact: same summation order, same rounding
        assert np.array_equal(d0.cpu().numpy(), acc)


def test_sparse_adam_matches_oracle(lib, T):
    from pkg import _native as N

    rng = np.random.default_rng(8)
    specs = [(5000, 16, 1), (300, 8, 2)]
    B = 700
    jobs, keep, host = _sparse_jobs(N, T, specs, B, rng, adam=True)
    ws = T.empty(int(lib.tt_sparse_workspace_bytes(len(specs), 2 * B, max(s[1] for s in specs))), dtype=T.uint8, device="cuda")
    lr_t = float(O.adam_lr_t(0.001, 3))
    N.check(lib.tt_sparse_sort(jobs, len(specs), ws.data_ptr(), ws.numel(), stream()))
    N.check(lib.tt_sparse_adam(jobs, len(specs), lr_t, 0.9, 0.999, 1e-7, ws.data_ptr(), ws.numel(), stream()))
    T.cuda.synchronize()
    for table, m, v, slices, dt, d0, d1 in host:
        O.adam_sparse(table, m, v, slices, 0.001, 3)
        # same formula, every op individually rounded on both sides; beta constants differ in the last ulp
        np.testing.assert_allclose(d0.cpu().numpy(), m, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(d1.cpu().numpy(), v, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(dt.cpu().numpy(), table, rtol=1e-5, atol=1e-6)


def test_dense_optimizers(lib, T):
    from pkg import _native as N

    rng = np.random.default_rng(9)
    n = 100_003
    w = rng.standard_normal(n).astype(np.float32); g = rng.standard_normal(n).astype(np.float32)
    acc = np.full(n, 0.1, np.float32)
    dw, dacc, dg = dev(T, w), dev(T, acc), dev(T, g)
    N.check(lib.tt_dense_adagrad(dw.data_ptr(), dacc.data_ptr(), dg.data_ptr(), n, 0.05, 1e-7, stream()))
    O.adagrad_dense(w, acc, g, 0.05)
    assert np.array_equal(dw.cpu().numpy(), w) and np.array_equal(dacc.cpu().numpy(), acc)   # bit-exact
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    dm, dv = dev(T, m), dev(T, v)
    N.check(lib.tt_dense_adam(dw.data_ptr(), dm.data_ptr(), dv.data_ptr(), dg.data_ptr(), n, float(O.adam_lr_t(0.001, 1)), 0.9, 0.999,
                              1e-7, stream()))
    O.adam_dense(w, m, v, g, 0.001, 1)
    np.testing.assert_allclose(dw.cpu().numpy(), w, rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------
def _index_call(lib, T, q, c, K, idx_base=0, impl=1):
    from pkg import _native as N

    dq, dc = dev(T, q), dev(T, c)
    nq, n, E = q.shape[0], c.shape[0], q.shape[1]
    s = T.empty((nq, K), dtype=T.float32, device="cuda"); i = T.empty((nq, K), dtype=T.int32, device="cuda")
    ws = T.empty(int(lib.tt_index_workspace_bytes(nq, n, E, K, impl, 0)), dtype=T.uint8, device="cuda")
    N.check(lib.tt_index_topk(dq.data_ptr(), E, dc.data_ptr(), E, None, None, nq, n, E, K, idx_base, s.data_ptr(), i.data_ptr(),
                              ws.data_ptr(), ws.numel(), impl, stream()))
    return s.cpu().numpy(), i.cpu().numpy()


def test_index_reference_fixture_with_oov_query(lib, T, golden):
    fix, _ = golden
    g = fix["brute_force"]
    rows = O.string_lookup(g["query_vocab"], g["queries"])
    q = np.array(g["query_table"], np.float32)[rows]
    c = np.array(g["candidate_embeddings"], np.float32)
    _, idx = _index_call(lib, T, q, c, g["k"])
    assert np.array(g["candidate_ids"])[idx].tolist() == g["expected"]      # tests/test_indices.py:119-129


def test_index_tie_break_kat_c(lib, T, golden):
    _, kat = golden
    s = np.array(kat["C"]["scores"], np.float32)          # scores as 1-d "embeddings" against unit queries
    q = np.ones((2, 1), np.float32)
    for row, want in zip(s, kat["C"]["indices"]):
        _, idx = _index_call(lib, T, q[:1], row.reshape(-1, 1), kat["C"]["k"])
        assert idx[0].tolist() == want


@pytest.mark.parametrize("nq,n,E,K", [(5, 1000, 32, 12), (70, 20_000, 64, 100), (64, 105_542, 64, 100), (3, 50, 16, 50),
                                      (9, 5000, 64, 1000), (1, 300, 32, 12), (130, 4097, 128, 128)])
def test_index_topk_bit_exact(lib, T, nq, n, E, K):
    rng = np.random.default_rng(10)
    q = np.maximum(rng.standard_normal((nq, E)) * 0.3, 0).astype(np.float32)
    c = (np.abs(rng.standard_normal((n, E))) * 0.1).astype(np.float32)
    want_s, want_i = O.index_topk(q, c, K, idx_base=1000)
    s, i = _index_call(lib, T, q, c, K, idx_base=1000)
    assert np.array_equal(i, want_i.astype(np.int32))
    assert np.array_equal(s, want_s)


def test_index_topk_dyadic_grid_many_ties(lib, T):
    rng = np.random.default_rng(11)
    q = (rng.integers(0, 5, size=(40, 32)) / 4.0).astype(np.float32)
    c = (rng.integers(0, 3, size=(9000, 32)) / 2.0).astype(np.float32)      # thousands of exactly equal scores
    want_s, want_i = O.index_topk(q, c, 100)
    s, i = _index_call(lib, T, q, c, 100)
    assert np.array_equal(i, want_i.astype(np.int32)) and np.array_equal(s, want_s)


def test_topk_merge_equals_unsharded(lib, T):
    from pkg import _native as N

    rng = np.random.default_rng(12)
    q = (rng.integers(0, 4, size=(33, 16)) / 2.0).astype(np.float32)
    c = (rng.integers(0, 4, size=(4000, 16)) / 2.0).astype(np.float32)
    K, G = 100, 4
    want_s, want_i = O.index_topk(q, c, K)
    per = 1000
    parts = [_index_call(lib, T, q, c[r * per:(r + 1) * per], K, idx_base=r * per) for r in range(G)]
    ps = dev(T, np.stack([p[0] for p in parts])); pi = dev(T, np.stack([p[1] for p in parts]))
    os_ = T.empty((33, K), dtype=T.float32, device="cuda"); oi = T.empty((33, K), dtype=T.int32, device="cuda")
    N.check(lib.tt_topk_merge(ps.data_ptr(), pi.data_ptr(), G, 33, K, os_.data_ptr(), oi.data_ptr(), stream()))
    assert np.array_equal(oi.cpu().numpy(), want_i.astype(np.int32)) and np.array_equal(os_.cpu().numpy(), want_s)


def test_recall_hits_integer_exact(lib, T):
    from pkg import _native as N

    rng = np.random.default_rng(13)
    nq, k = 3000, 100
    cand = rng.integers(0, 500, size=(nq, k)).astype(np.int32)
    truth = rng.integers(-1, 500, size=nq).astype(np.int32)
    ks = np.array([1, 10, 100], np.int32)
    want = np.zeros(3, np.int32)
    O.c_lib().tto_recall_hits(cand.astype(np.int64).ctypes.data_as(ctypes.c_void_p), truth.astype(np.int64).ctypes.data_as(ctypes.c_void_p),
                              nq, k, ks.ctypes.data_as(ctypes.c_void_p), 3, want.ctypes.data_as(ctypes.c_void_p))
    hits = T.zeros(3, dtype=T.int32, device="cuda")
    dcand, dtruth = dev(T, cand), dev(T, truth)
    for _ in range(2):                                     # accumulates across calls
        N.check(lib.tt_recall_hits(dcand.data_ptr(), k, dtruth.data_ptr(), nq, ks.ctypes.data, 3, hits.data_ptr(), stream()))
    assert hits.cpu().numpy().tolist() == (2 * want).tolist()


# ---------------------------------------------------------------------------------------------------
# row-sharded tables (tt_feature.shards, tt_sparse_job.shard_rank/shard_world).  On one GPU every "shard" is local
# memory; across processes the same kernels dereference peer-mapped pointers (tests/test_gpu_multi.py).
# ---------------------------------------------------------------------------------------------------
def _shard(T, full, g):
    rows, e = full.shape
    local = (rows + g - 1) // g
    out = []
    for r in range(g):
        s = np.zeros((local, e), np.float32)
        part = full[r::g]
        s[: part.shape[0]] = part
        out.append(dev(T, s))
    return out


@pytest.mark.parametrize("g", [2, 3, 8])
def test_row_sharded_gather_and_sparse_adagrad_equal_unsharded(lib, T, g):
    from pkg import _native as N

    rng = np.random.default_rng(40 + g)
    rows, e, B = 1001, 64, 700
    table = rng.standard_normal((rows, e)).astype(np.float32)
    ids = np.minimum(rng.zipf(1.2, size=B) - 1, rows - 1).astype(np.int32)
    ids[5] = rows + 3                                        # out of range -> OOV row 0 (owned by shard 0)
    d_ids = dev(T, ids)
    shards = _shard(T, table, g)
    ptrs = T.tensor([s.data_ptr() for s in shards], dtype=T.int64, device="cuda")
    # gather: sharded descriptor == plain descriptor
    d_table = dev(T, table)
    X0 = T.empty((B, e), dtype=T.float32, device="cuda"); X1 = T.empty_like(X0)
    N.check(lib.tt_gather_concat(feats_array(N, [dict(table=d_table.data_ptr(), src=d_ids.data_ptr(), rows=rows, e=e, col=0)]), 1, B, e,
                                 X0.data_ptr(), e, stream()))
    N.check(lib.tt_gather_concat(feats_array(N, [dict(table=ptrs.data_ptr(), src=d_ids.data_ptr(), rows=rows, e=e, col=0, shards=g)]), 1, B, e,
                                 X1.data_ptr(), e, stream()))
    assert T.equal(X0, X1)
    # sparse Adagrad: every "rank" applies its own rows; reassembled == the oracle on the whole table
    grad = rng.standard_normal((B, e)).astype(np.float32)
    d_grad = dev(T, grad)
    accs = [T.full_like(s, 0.1) for s in shards]
    ws = T.empty(int(lib.tt_sparse_workspace_bytes(1, B, e)), dtype=T.uint8, device="cuda")
    for r in range(g):
        jobs = (N.TTSparseJob * 1)()
        jobs[0].table, jobs[0].slot0, jobs[0].slot1 = shards[r].data_ptr(), accs[r].data_ptr(), None
        jobs[0].rows, jobs[0].e, jobs[0].nsrc, jobs[0].n_per_src = rows, e, 1, B
        jobs[0].shard_rank, jobs[0].shard_world = r, g
        jobs[0].ids[0], jobs[0].grad[0], jobs[0].grad_ld[0] = d_ids.data_ptr(), d_grad.data_ptr(), e
        N.check(lib.tt_sparse_sort(jobs, 1, ws.data_ptr(), ws.numel(), stream()))
        N.check(lib.tt_sparse_adagrad(jobs, 1, 0.05, 1e-7, ws.data_ptr(), ws.numel(), stream()))
    T.cuda.synchronize()
    # oracle: the update of shard r is the de-duplicated update of the entries it owns, in their original order (runs are summed
    # piecewise along 32-entry blocks of the SORTED OWNED list, so the rounding of a heavily duplicated row may differ in the last
    # bit from the unsharded update -- compared with a tolerance below, bit-exactly against the per-shard oracle here)
    ids_c = np.where((ids < 0) | (ids >= rows), 0, ids)
    whole = table.copy(); whole_acc = np.full_like(table, 0.1)
    O.adagrad_sparse(whole, whole_acc, O.IndexedSlices(ids_c, grad), 0.05)
    for r in range(g):
        n_r = (rows - r + g - 1) // g
        mine = ids_c % g == r
        t_r = np.ascontiguousarray(table[r::g]); a_r = np.full_like(t_r, 0.1)
        O.adagrad_sparse(t_r, a_r, O.IndexedSlices(ids_c[mine] // g, grad[mine]), 0.05)
        assert np.array_equal(shards[r].cpu().numpy()[:n_r], t_r)
        assert np.array_equal(accs[r].cpu().numpy()[:n_r], a_r)
        np.testing.assert_allclose(t_r, whole[r::g], rtol=0, atol=1e-6)


def test_row_sharded_sparse_adam_equals_per_shard_oracle(lib, T):
    """Legacy (non-lazy) Adam on a row-sharded table: whole-shard decay sweep, scatter of the owned entries, whole-shard update."""
    from pkg import _native as N

    rng = np.random.default_rng(77)
    g, rows, e, B = 3, 400, 16, 300
    table = rng.standard_normal((rows, e)).astype(np.float32)
    m0 = (rng.standard_normal((rows, e)) * 0.01).astype(np.float32); v0 = (rng.random((rows, e)) * 0.01).astype(np.float32)
    ids = np.minimum(rng.zipf(1.3, size=B) - 1, rows - 1).astype(np.int32)
    grad = rng.standard_normal((B, e)).astype(np.float32)
    d_ids, d_grad = dev(T, ids), dev(T, grad)
    ws = T.empty(int(lib.tt_sparse_workspace_bytes(1, B, e)), dtype=T.uint8, device="cuda")
    lr_t = float(O.adam_lr_t(0.01, 3))
    for r in range(g):
        sh = [np.ascontiguousarray(a[r::g]) for a in (table, m0, v0)]
        local = (rows + g - 1) // g
        pad = [np.zeros((local, e), np.float32) for _ in range(3)]
        for p_, s_ in zip(pad, sh):
            p_[: s_.shape[0]] = s_
        dt, dm, dv = (dev(T, a) for a in pad)
        jobs = (N.TTSparseJob * 1)()
        jobs[0].table, jobs[0].slot0, jobs[0].slot1 = dt.data_ptr(), dm.data_ptr(), dv.data_ptr()
        jobs[0].rows, jobs[0].e, jobs[0].nsrc, jobs[0].n_per_src = rows, e, 1, B
        jobs[0].shard_rank, jobs[0].shard_world = r, g
        jobs[0].ids[0], jobs[0].grad[0], jobs[0].grad_ld[0] = d_ids.data_ptr(), d_grad.data_ptr(), e
        N.check(lib.tt_sparse_sort(jobs, 1, ws.data_ptr(), ws.numel(), stream()))
        N.check(lib.tt_sparse_adam(jobs, 1, lr_t, 0.9, 0.999, 1e-7, ws.data_ptr(), ws.numel(), stream()))
        T.cuda.synchronize()
        mine = ids % g == r
        t_r, m_r, v_r = (a.copy() for a in sh)
        O.adam_sparse(t_r, m_r, v_r, O.IndexedSlices(ids[mine] // g, grad[mine]), 0.01, 3)
        n_r = t_r.shape[0]
        np.testing.assert_allclose(dt.cpu().numpy()[:n_r], t_r, rtol=1e-5, atol=1e-6)   # tolerances of the unsharded Adam test
        np.testing.assert_allclose(dm.cpu().numpy()[:n_r], m_r, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(dv.cpu().numpy()[:n_r], v_r, rtol=1e-5, atol=1e-9)


def test_stage_columns_one_launch_device_pinned_and_int64(lib):
    """tt_stage_columns: device and pinned-host sources, int64 narrowing, unaligned ends and row counts that are not multiples of 4."""
    import torch

    from pkg import _native as N

    for rows in (1, 7, 1000, 8193):
        g = torch.Generator().manual_seed(rows)
        ids_dev = torch.randint(0, 1 << 30, (rows,), generator=g, dtype=torch.int32).cuda()
        f_pin = torch.rand((rows,), generator=g).pin_memory()
        i64_pin = torch.randint(0, 1 << 30, (rows,), generator=g, dtype=torch.int64).pin_memory()
        odd = torch.randint(0, 99, (rows + 1,), generator=g, dtype=torch.int32).cuda()[1:]        # 4-byte aligned only
        outs = [torch.full((rows + 4,), -7, dtype=torch.int32, device="cuda") for _ in range(4)]
        fout = torch.full((rows + 4,), -7.0, device="cuda")
        cols = (N.TTStageCol * 4)()
        for i, (src, dst, kind) in enumerate(((ids_dev, outs[0], 0), (f_pin, fout, 0), (i64_pin, outs[2], 1), (odd, outs[3][1:], 0))):
            cols[i].src, cols[i].dst, cols[i].kind = src.data_ptr(), dst.data_ptr(), kind
        c0 = lib.tt_launch_count()
        N.check(lib.tt_stage_columns(cols, 4, rows, N.stream_ptr()))
        assert lib.tt_launch_count() - c0 == 1
        torch.cuda.synchronize()
        assert torch.equal(outs[0][:rows], ids_dev) and bool((outs[0][rows:] == -7).all())
        assert torch.equal(fout[:rows].cpu(), f_pin) and bool((fout[rows:] == -7).all())
        assert torch.equal(outs[2][:rows].cpu(), i64_pin.to(torch.int32)) and bool((outs[2][rows:] == -7).all())
        assert torch.equal(outs[3][1:rows + 1], odd) and int(outs[3][0]) == -7 and bool((outs[3][rows + 1:] == -7).all())
    assert lib.tt_stage_columns(None, 2, 5, None) != 0
    assert lib.tt_stage_columns(cols, N.TT_MAX_STAGE_COLS + 1, 5, None) != 0
