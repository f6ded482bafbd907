"""2-GPU NCCL tests of data-parallel training and the sharded index (skipped with fewer than 2 GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _features():
    from pkg.schema import dtypes as tt
    from pkg.schema.features import Feature, FeatureFamily

    qf = [Feature("age", tt.float32, FeatureFamily.QUERY), Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=32)]
    cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=32),
          Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=8)]
    qf[1].set_vocab_size(600); cf[0].set_vocab_size(300); cf[1].set_vocab_size(50)
    return qf, cf


def _batch(rng, b):
    art = np.minimum(rng.zipf(1.3, size=b), 300).astype(np.int32)
    return {"age": rng.random((b, 1)).astype(np.float32), "customer_id": rng.integers(0, 601, size=(b, 1)).astype(np.int32),
            "article_id": art.reshape(b, 1), "colour_group_name": (art % 50 + 1).reshape(b, 1).astype(np.int32)}


def _worker(rank, world, port, out, shard_tables, peer_sync, global_negatives=False, impl=1, opt="adagrad"):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
    import torch
    import torch.distributed as dist

    from pkg import _native as N
    from pkg.modelling._device import set_seed
    from pkg.modelling.distributed import DataParallel, make_sharded_index
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        set_seed(17)
        qf, cf = _features()
        model = TwoTowerModel(qf, cf, "article_id", 32, candidate_prob_lookup={str(i + 1): 1.0 / 300 for i in range(300)})
        model.impl = impl
        model.compile(optimizer=OptimizerFactory.get_optimizer(opt, {"learning_rate": 0.05 if opt == "adagrad" else 0.01}))
        before = {k: v.copy() for k, v in model.state_arrays().items()}
        dp = DataParallel(model, shard_tables=shard_tables, peer_sync=peer_sync, global_negatives=global_negatives)
        assert dp.shard_tables == shard_tables and dp.peer_sync == peer_sync
        rng = np.random.default_rng(100)
        batches = [_batch(rng, 96) for _ in range(world)]
        loss = float(model.train_step(batches[rank])["loss"])
        after = model.state_arrays()
        # two more steps through the captured-graph path (step 2 captures, step 3 replays): cross-step ordering of the peer reads
        model.use_cuda_graph = True
        model._steps.clear()
        more = [_batch(rng, 96) for _ in range(3 * world)]
        for k in range(3):
            model.train_step(more[k * world + rank])
        dp.barrier()
        after3 = model.state_arrays()
        # sharded index over the (now trained) candidate tower
        art = np.arange(1, 301, dtype=np.int32)
        emb = model.candidate_tower({"article_id": art.reshape(-1, 1), "colour_group_name": (art % 50 + 1).reshape(-1, 1)})
        index = make_sharded_index(10, model.query_tower, [(art, emb)])
        q = {"age": np.linspace(0, 1, 40, dtype=np.float32).reshape(-1, 1), "customer_id": np.arange(40, dtype=np.int32).reshape(-1, 1)}
        ids = index(q)
        q_emb = model.query_tower(q).cpu().numpy()
        np.savez(out, loss=loss, ids=ids, q_emb=q_emb, c_emb=emb.cpu().numpy(), **{"before/" + k: v for k, v in before.items()},
                 **{"after/" + k: v for k, v in after.items()}, **{"after3/" + k: v for k, v in after3.items()})
    finally:
        dist.destroy_process_group()


def _run(tmp_path, world, shard_tables, peer_sync=False, global_negatives=False, impl=1, opt="adagrad"):
    import torch.multiprocessing as mp

    port = _free_port()
    outs = [str(tmp_path / f"r{r}_{int(shard_tables)}{int(peer_sync)}{int(global_negatives)}{impl}{opt}.npz") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, outs[r], shard_tables, peer_sync, global_negatives, impl, opt)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(280)
        assert p.exitcode == 0
    return [np.load(o) for o in outs]


@pytest.mark.timeout(600)
def test_data_parallel_step_and_sharded_index_on_two_gpus(tmp_path):
    """Replicated tables (all-gathered gradient rows) and row-sharded tables (rows gathered / gradients pulled over NVLink peer
    memory) must agree bit for bit with each other, and with the oracle's G-batches-summed step."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import two_tower_oracle as O

    world = 2
    res = _run(tmp_path, world, False)
    res_sh = _run(tmp_path, world, True)                      # row-sharded tables, ranks ordered by the two small NCCL all-gathers
    res_pe = _run(tmp_path, world, True, True)                # row-sharded tables, device barriers + peer reads only (one graph per step)
    for k in res_sh[0].files:                                 # same arithmetic, different synchronisation: bit-identical
        for r in range(world):
            assert np.array_equal(res_sh[r][k], res_pe[r][k]), k
    # every saved array (losses, embeddings, weights after 1 and after 4 steps).  The two layouts sum a duplicated row's gradients
    # along different 32-entry block cuts of their sorted id lists, so heavily duplicated rows may differ in the last bits
    # (tests/test_gpu_kernels.py pins each layout bit-exactly to the oracle); everything else is identical.
    for k in res[0].files:
        for r in range(world):
            a, b_ = res[r][k], res_sh[r][k]
            if a.dtype.kind == "f":
                np.testing.assert_allclose(b_, a, rtol=0, atol=1e-5 * (1.0 + float(np.abs(a).max())), err_msg=k)   # the dense-gradient sums round differently (torch.sum vs rank-order peer sum); Adam's m / sqrt(v) amplifies an ulp
            else:
                assert np.array_equal(a, b_), k
    keys = [k for k in res[0].files if k.startswith("after")]
    for k in keys:                                     # replicas stay bit-identical
        assert np.array_equal(res[0][k], res[1][k]), k
    assert np.array_equal(res[0]["ids"], res[1]["ids"])
    # oracle: G batches at the same weights, gradients summed, one Adagrad apply
    b = {k[len("before/"):]: res[0][k] for k in res[0].files if k.startswith("before/")}
    qt = O.OracleTower([O.OracleFeature("age", False), O.OracleFeature("customer_id", True, 32)],
                       {"customer_id": b["query_tower/embedding/customer_id"].copy()},
                       [(b["query_tower/dense_0/kernel"].copy(), b["query_tower/dense_0/bias"].copy())])
    ct = O.OracleTower([O.OracleFeature("article_id", True, 32), O.OracleFeature("colour_group_name", True, 8)],
                       {"article_id": b["candidate_tower/embedding/article_id"].copy(), "colour_group_name": b["candidate_tower/embedding/colour_group_name"].copy()},
                       [(b["candidate_tower/dense_0/kernel"].copy(), b["candidate_tower/dense_0/bias"].copy())])
    rng = np.random.default_rng(100)
    batches = [_batch(rng, 96) for _ in range(world)]
    p_rows = np.ones(301, np.float32); p_rows[1:] = np.float32(1.0 / 300)
    gs = [O.train_step_grads(qt, ct, {"customer_id": bt["customer_id"]}, {"age": bt["age"]},
                             {"article_id": bt["article_id"], "colour_group_name": bt["colour_group_name"]}, {},
                             p_rows[bt["article_id"].reshape(-1)]) for bt in batches]
    for r in range(world):
        assert abs(float(res[r]["loss"]) - gs[r].loss) <= 1e-5 * abs(gs[r].loss)
    dw = sum(g.dense_c[0][0] for g in gs); w = ct.dense[0][0].copy(); aw = np.full_like(w, 0.1)
    O.adagrad_dense(w, aw, dw, 0.05)
    np.testing.assert_allclose(res[0]["after/candidate_tower/dense_0/kernel"], w, rtol=0, atol=2e-4)
    ids = np.concatenate([g.tables_c["article_id"].indices for g in gs]); rows = np.concatenate([g.tables_c["article_id"].values for g in gs])
    t = ct.tables["article_id"].copy(); acc = np.full_like(t, 0.1)
    O.adagrad_sparse(t, acc, O.IndexedSlices(ids, rows), 0.05)
    np.testing.assert_allclose(res[0]["after/candidate_tower/embedding/article_id"], t, rtol=0, atol=2e-4)
    # sharded index == unsharded oracle, bit-exact ids
    _, want = O.index_topk(res[0]["q_emb"], res[0]["c_emb"], 10)
    assert np.array_equal(res[0]["ids"], np.arange(1, 301, dtype=np.int32)[want])


@pytest.mark.timeout(600)
@pytest.mark.parametrize("impl", [1, 0])          # exact CUDA-core contraction; tensor-core path (TT_IMPL_AUTO)
def test_global_negatives_equal_one_process_on_the_concatenated_batch(tmp_path, impl):
    """BASELINE configs[4] / SURVEY.md 8e: with all-gathered candidates every rank scores its B rows against the G.B candidates
    of all ranks; summed over ranks this is exactly the reference's train_step on the concatenated batch."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import two_tower_oracle as O

    world = 2
    res = _run(tmp_path, world, True, True, True, impl)
    b = {k[len("before/"):]: res[0][k] for k in res[0].files if k.startswith("before/")}
    qt = O.OracleTower([O.OracleFeature("age", False), O.OracleFeature("customer_id", True, 32)],
                       {"customer_id": b["query_tower/embedding/customer_id"].copy()},
                       [(b["query_tower/dense_0/kernel"].copy(), b["query_tower/dense_0/bias"].copy())])
    ct = O.OracleTower([O.OracleFeature("article_id", True, 32), O.OracleFeature("colour_group_name", True, 8)],
                       {"article_id": b["candidate_tower/embedding/article_id"].copy(), "colour_group_name": b["candidate_tower/embedding/colour_group_name"].copy()},
                       [(b["candidate_tower/dense_0/kernel"].copy(), b["candidate_tower/dense_0/bias"].copy())])
    rng = np.random.default_rng(100)
    batches = [_batch(rng, 96) for _ in range(world)]
    cat = {k: np.concatenate([bt[k] for bt in batches], axis=0) for k in batches[0]}
    p_rows = np.ones(301, np.float32); p_rows[1:] = np.float32(1.0 / 300)
    g = O.train_step_grads(qt, ct, {"customer_id": cat["customer_id"]}, {"age": cat["age"]},
                           {"article_id": cat["article_id"], "colour_group_name": cat["colour_group_name"]}, {},
                           p_rows[cat["article_id"].reshape(-1)])
    tol = 1e-5 if impl == 1 else 1e-3                           # north star: 1e-3 relative on the TF32/fp16 path
    total = sum(float(res[r]["loss"]) for r in range(world))   # every rank reports the loss of its own rows
    assert abs(total - g.loss) <= tol * abs(g.loss)
    atol = 2e-4 if impl == 1 else 2e-3
    for r in range(world):                                      # replicas / shards agree on the assembled state
        for k in res[0].files:
            if k.startswith("after/"):
                assert np.array_equal(res[0][k], res[r][k]), k
    w = ct.dense[0][0].copy(); aw = np.full_like(w, 0.1)
    O.adagrad_dense(w, aw, g.dense_c[0][0], 0.05)
    np.testing.assert_allclose(res[0]["after/candidate_tower/dense_0/kernel"], w, rtol=0, atol=atol)
    wq = qt.dense[0][0].copy(); awq = np.full_like(wq, 0.1)
    O.adagrad_dense(wq, awq, g.dense_q[0][0], 0.05)
    np.testing.assert_allclose(res[0]["after/query_tower/dense_0/kernel"], wq, rtol=0, atol=atol)
    for name, tower, slices in (("candidate_tower/embedding/article_id", ct, g.tables_c["article_id"]),
                                ("query_tower/embedding/customer_id", qt, g.tables_q["customer_id"])):
        t = tower.tables[name.split("/")[-1]].copy(); acc = np.full_like(t, 0.1)
        O.adagrad_sparse(t, acc, slices, 0.05)
        np.testing.assert_allclose(res[0]["after/" + name], t, rtol=0, atol=atol)


@pytest.mark.timeout(600)
def test_data_parallel_adam_row_sharded_equals_replicated(tmp_path):
    """Legacy (non-lazy) Adam: whole-table sweeps run on each rank's shard; four steps agree with the replicated layout."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    rep = _run(tmp_path, world, False, opt="adam")
    sh = _run(tmp_path, world, True, True, opt="adam")
    for k in rep[0].files:
        for r in range(world):
            a, b_ = rep[r][k], sh[r][k]
            if a.dtype.kind == "f":
                np.testing.assert_allclose(b_, a, rtol=0, atol=1e-5 * (1.0 + float(np.abs(a).max())), err_msg=k)   # the dense-gradient sums round differently (torch.sum vs rank-order peer sum); Adam's m / sqrt(v) amplifies an ulp
            else:
                assert np.array_equal(a, b_), k
    moved = [k for k in rep[0].files if k.startswith("after3/") and "embedding" in k]
    assert moved and all(not np.array_equal(rep[0][k], rep[0]["before/" + k[len("after3/"):]]) for k in moved)


def _worker_owner_init(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
    import torch
    import torch.distributed as dist

    from pkg.modelling._device import set_seed
    from pkg.modelling.distributed import DataParallel
    from pkg.modelling.models.two_tower_model import TwoTowerModel
    from pkg.modelling.optimizer_factory import OptimizerFactory

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        def build(seed):
            set_seed(seed)
            qf, cf = _features()
            m = TwoTowerModel(qf, cf, "article_id", 32, candidate_prob_lookup={str(i + 1): 1.0 / 300 for i in range(300)})
            m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
            return m

        a = build(17 + rank)                              # different seeds per rank: rank 0's must win, as with the broadcast
        assert not any(t.materialised for _, _, t in a._tables())
        DataParallel(a, shard_tables=True)
        assert all(t.materialised and t.weight.shape[0] == t.local_rows for _, _, t in a._tables())   # only the shard exists
        sa = a.state_arrays()
        b = build(17)                                     # the single-process model of rank 0's seed
        sb = b.state_arrays()
        c = build(17 + rank)
        for _, _, t in c._tables():
            t.weight                                      # materialise first: the broadcast-and-slice path
        DataParallel(c, shard_tables=True)
        rng = np.random.default_rng(100)
        batches = [_batch(rng, 96) for _ in range(world)]
        la, lc = float(a.train_step(batches[rank])["loss"]), float(c.train_step(batches[rank])["loss"])
        a.dist.barrier(); c.dist.barrier()
        ta, tc = a.state_arrays(), c.state_arrays()
        # five batch shapes on one model: the oldest step workspace is evicted and its peer-shared buffers are closed collectively
        # (PeerBuffer.close); the evicted shape is rebuilt when it comes back
        for bs in (64, 48, 32, 16):
            more = float(c.train_step(_batch(rng, bs))["loss"])
            assert np.isfinite(more)
        again = float(c.train_step(batches[rank])["loss"])
        assert np.isfinite(again) and len(c._steps) == 4
        np.savez(out, la=la, lc=lc, **{"a/" + k: v for k, v in sa.items()}, **{"b/" + k: v for k, v in sb.items()},
                 **{"ta/" + k: v for k, v in ta.items()}, **{"tc/" + k: v for k, v in tc.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_row_shards_initialised_by_their_owners_equal_the_unsharded_table(tmp_path):
    """A table that is row-sharded before its first use is filled shard by shard on the owners (no whole-table allocation or
    broadcast); it must hold the single-process table of rank 0's seed, and training must not depend on which path built it."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    outs = [str(tmp_path / f"owner{r}.npz") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker_owner_init, args=(r, world, port, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(280)
        assert p.exitcode == 0
    res = [np.load(o) for o in outs]
    keys = [k[2:] for k in res[0].files if k.startswith("a/")]
    assert any("embedding" in k for k in keys)
    for r in res:
        for k in keys:
            assert np.array_equal(r["a/" + k], r["b/" + k]), k              # owner-initialised == unsharded, bit for bit
            assert np.array_equal(r["ta/" + k], r["tc/" + k]), k            # and trains like the broadcast-and-slice path
        assert float(r["la"]) == float(r["lc"])
    for k in keys:
        assert np.array_equal(res[0]["ta/" + k], res[1]["ta/" + k]), k
