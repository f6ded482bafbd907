"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing in pkg.modelling.distributed: the collective helpers
are the product's own; the arithmetic around them is the oracle's (no GPU here).  They check the two exchange
protocols of SURVEY.md 8(e): data-parallel training == G batches' gradients summed + one optimizer apply, and
sharded index + all-gather merge == unsharded index."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _towers(rng):
    from oracle import two_tower_oracle as O

    qt = O.OracleTower([O.OracleFeature("q", True, 8)], {"q": rng.standard_normal((50, 8)).astype(np.float32) * 0.1},
                       [(rng.standard_normal((8, 4)).astype(np.float32) * 0.3, np.zeros(4, np.float32))])
    ct = O.OracleTower([O.OracleFeature("c", True, 8)], {"c": rng.standard_normal((20, 8)).astype(np.float32) * 0.1},
                       [(rng.standard_normal((8, 4)).astype(np.float32) * 0.3, np.zeros(4, np.float32))])
    return qt, ct


def _batches(rng, g, b):
    return [({"q": rng.integers(0, 50, size=b)}, {"c": rng.integers(0, 20, size=b)}) for _ in range(g)]


def _dp_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
    from oracle import two_tower_oracle as O
    from pkg.modelling.distributed import allgather_into, allreduce_sum_, shard_bounds

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        qt, ct = _towers(rng)
        batches = _batches(rng, world, 16)
        q_ids, c_ids = batches[rank]
        g = O.train_step_grads(qt, ct, q_ids, {}, c_ids, {}, None)
        # dense: one flat SUM all-reduce
        flat = torch.from_numpy(np.concatenate([g.dense_c[0][0].ravel(), g.dense_c[0][1].ravel()]).copy())
        allreduce_sum_(flat)
        # sparse: all-gather (ids, gradient rows), rank-major
        ids_all = torch.zeros(world * 16, dtype=torch.int64)
        rows_all = torch.zeros(world * 16, 8)
        allgather_into(ids_all, torch.from_numpy(g.tables_c["c"].indices.astype(np.int64)))
        allgather_into(rows_all, torch.from_numpy(g.tables_c["c"].values.copy()))
        table = ct.tables["c"].copy(); acc = np.full_like(table, 0.1)
        O.adagrad_sparse(table, acc, O.IndexedSlices(ids_all.numpy(), rows_all.numpy()), 0.05)
        # row-sharded table (DataParallel(shard_tables=True)): row i is owned by rank i % G at local row i // G; the owner applies
        # the de-duplicated update of the entries it owns, in (rank, position) order; shards are all-gathered to compare
        mine = ids_all.numpy() % world == rank
        local = (20 + world - 1) // world
        shard = np.zeros((local, 8), np.float32); shard_acc = np.full_like(shard, 0.1)
        part = ct.tables["c"][rank::world]
        shard[: part.shape[0]] = part
        O.adagrad_sparse(shard, shard_acc, O.IndexedSlices(ids_all.numpy()[mine] // world, rows_all.numpy()[mine]), 0.05)
        shards = torch.zeros(world * local, 8)
        allgather_into(shards, torch.from_numpy(shard))
        table_sharded = np.zeros_like(table)
        for r in range(world):
            n_r = (20 - r + world - 1) // world
            table_sharded[r::world] = shards.numpy().reshape(world, local, 8)[r, :n_r]
        # sharded index: local exact top-k with global indices, all-gather, merge
        corpus = np.random.default_rng(5).integers(0, 4, size=(103, 6)).astype(np.float32)
        queries = np.random.default_rng(6).integers(0, 4, size=(9, 6)).astype(np.float32)
        lo, hi = shard_bounds(103, rank, world)
        s, i = O.index_topk(queries, corpus[lo:hi], 10, idx_base=lo)
        all_s = torch.zeros(world * 9, 10); all_i = torch.zeros(world * 9, 10, dtype=torch.int64)
        allgather_into(all_s, torch.from_numpy(s)); allgather_into(all_i, torch.from_numpy(i))
        ms, mi = O.merge_topk(all_s.numpy().reshape(world, 9, 10), all_i.numpy().reshape(world, 9, 10), 10)
        np.savez(out, flat=flat.numpy(), table=table, acc=acc, ms=ms, mi=mi, table_sharded=table_sharded)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dp_and_sharded_index_protocols_world2(tmp_path):
    import torch.multiprocessing as mp

    from oracle import two_tower_oracle as O

    world, port = 2, _free_port()
    outs = [str(tmp_path / f"r{r}.npz") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    res = [np.load(o) for o in outs]
    # replicas agree bit for bit
    # the row-sharded layout reproduces the replicated update (last-bit differences only where a row's duplicates straddle
    # different 32-entry blocks of the two sorted lists; none at this size)
    np.testing.assert_allclose(res[0]["table_sharded"], res[0]["table"], rtol=0, atol=1e-7)
    for k in ("flat", "table", "acc", "ms", "mi", "table_sharded"):
        assert np.array_equal(res[0][k], res[1][k]), k
    # == the single-process statement: G batches at the same weights, gradients summed, one apply
    rng = np.random.default_rng(0)
    qt, ct = _towers(rng)
    batches = _batches(rng, world, 16)
    gs = [O.train_step_grads(qt, ct, q, {}, c, {}, None) for q, c in batches]
    want_flat = sum(np.concatenate([g.dense_c[0][0].ravel(), g.dense_c[0][1].ravel()]) for g in gs)
    np.testing.assert_allclose(res[0]["flat"], want_flat, rtol=1e-6, atol=1e-7)
    ids = np.concatenate([g.tables_c["c"].indices for g in gs]); rows = np.concatenate([g.tables_c["c"].values for g in gs])
    table = ct.tables["c"].copy(); acc = np.full_like(table, 0.1)
    O.adagrad_sparse(table, acc, O.IndexedSlices(ids, rows), 0.05)
    assert np.array_equal(res[0]["table"], table) and np.array_equal(res[0]["acc"], acc)
    # sharded index == unsharded (many exact ties in this integer corpus: the tie rule must survive the merge)
    corpus = np.random.default_rng(5).integers(0, 4, size=(103, 6)).astype(np.float32)
    queries = np.random.default_rng(6).integers(0, 4, size=(9, 6)).astype(np.float32)
    s, i = O.index_topk(queries, corpus, 10)
    assert np.array_equal(res[0]["mi"], i) and np.array_equal(res[0]["ms"], s)


def _recall_worker(rank, world, port, out):
    import torch.distributed as dist

    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
    from pkg.modelling.metrics.index_recall import IndexRecall

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        class Fixed:                       # an index returning the same 5 candidates for every query (host path of IndexRecall)
            def __call__(self, queries):
                n = len(queries["q"])
                return np.tile(np.array([["a", "b", "c", "d", "e"]], dtype=object), (n, 1))

        truth = np.array(list("abcdxaybzc"), dtype=object)            # the reference's ragged-batch fixture, split over the ranks
        metric = IndexRecall(Fixed(), ks=[1, 2, 5])
        mine = truth[rank::world]
        for lo in range(0, len(mine), 2):                             # ragged batches
            t = mine[lo:lo + 2]
            metric({"q": np.zeros((len(t), 1))}, t.reshape(-1, 1))
        metric.all_reduce()
        metric.all_reduce()                                           # idempotent: the ranks' own counts are what is summed
        np.savez(out, hits=np.array([metric.hits[k] for k in (1, 2, 5)]), seen=metric.seen, m=np.array([metric.metric[k] for k in (1, 2, 5)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_recall_counters_all_reduce_world2(tmp_path):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    outs = [str(tmp_path / f"rec{r}.npz") for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_recall_worker, args=(r, world, port, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    res = [np.load(o) for o in outs]
    truth = np.array(list("abcdxaybzc"), dtype=object)
    cand = np.array(list("abcde"), dtype=object)
    want_hits = [int(sum(t in cand[:k] for t in truth)) for k in (1, 2, 5)]
    for r in res:
        assert r["hits"].tolist() == want_hits and int(r["seen"]) == 10
        assert r["hits"].dtype == np.int32 and r["m"].dtype == np.float64
        assert r["m"].tolist() == [h / 10 for h in want_hits]


def test_shard_bounds_cover_everything():
    from pkg.modelling.distributed import shard_bounds

    for n in (0, 1, 7, 105_542):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
