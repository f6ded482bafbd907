#!/usr/bin/env python
"""Development diagnostic: train the bench model into the popularity-collapsed regime (large SUM-loss batches), then count how
many queries of an index batch take the exact fallback and why (listed columns per query vs the list capacity, hit-log fill)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from pkg import _native as N  # noqa: E402
from pkg.modelling.indices.brute_force import BruteForceIndex  # noqa: E402

lib = N.load()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
model = bench.build_gpu_model()
rng = np.random.default_rng(1)
batches = [{k: torch.from_numpy(v).cuda() for k, v in bench.make_batch(rng, B).items()} for _ in range(4)]
art = np.arange(1, bench.V_ARTICLES + 1, dtype=np.int32)
x = {"article_id": art.reshape(-1, 1), "product_type_name": (art % bench.V_PTYPE + 1).reshape(-1, 1), "colour_group_name": (art % bench.V_COLOUR + 1).reshape(-1, 1)}
done = 0
for upto in (0, steps // 4, steps):
    while done < upto:
        model.train_step(batches[done % 4]); done += 1
    torch.cuda.synchronize()
    emb = model.candidate_tower(x)
    index = BruteForceIndex(100, model.query_tower, [(art, emb)])
    norms = emb.norm(dim=1)
    for seed in range(3):
        r = np.random.default_rng(77 + seed)
        q = {"age": r.random((2048, 1)).astype(np.float32), "customer_id": r.integers(1, bench.V_CUSTOMERS + 1, size=(2048, 1)).astype(np.int32)}
        qe = index._embed_queries(q)
        for _ in range(2):
            s, i = index.search(qe)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); index.search(qe); ev1.record(); torch.cuda.synchronize()
        out8 = (ctypes.c_int64 * 8)()
        lib.tt_debug_index_layout(2048, bench.V_ARTICLES, 64, 100, 1, out8)
        ws = index._ws
        flags = ws[out8[0]:out8[0] + 4 * 2048].view(torch.int32)
        ccnt = ws[out8[1]:out8[1] + 4 * 2048].view(torch.int32)
        logs = ws[out8[2]:out8[2] + 4 * out8[5]].view(torch.int32)
        gap = (s[:, 0] - s[:, 99]) / (qe.norm(dim=1) * norms.max())
        print(f"steps {done:5d} seed {seed}: search {ev0.elapsed_time(ev1):.3f} ms  flagged {int(flags.sum())}/2048  listed per query: median {int(ccnt.median())} "
              f"p99 {int(ccnt.float().quantile(0.99))} max {int(ccnt.max())} (cap {out8[3]})  log fill max {int(logs.max())} (cap {out8[4]})  "
              f"(s1 - s100)/(|q| max|c|): median {float(gap.median()):.2e}  |c| max {float(norms.max()):.2f} median {float(norms.median()):.2f}", flush=True)
