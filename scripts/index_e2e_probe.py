"""Where does the host-facing index call spend its time? (development tool, GPU only)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import torch
import bench
from pkg.modelling.indices.brute_force import BruteForceIndex

model = bench.build_gpu_model()
art = np.arange(1, bench.V_ARTICLES + 1, dtype=np.int32)
pairs = []
for lo in range(0, bench.V_ARTICLES, 10000):
    a = art[lo:lo + 10000]
    x = {"article_id": a.reshape(-1, 1), "product_type_name": (a % bench.V_PTYPE + 1).reshape(-1, 1), "colour_group_name": (a % bench.V_COLOUR + 1).reshape(-1, 1)}
    pairs.append((a, model.candidate_tower(x)))
index = BruteForceIndex(100, model.query_tower, pairs)
rng = np.random.default_rng(0)
h = {"age": rng.random((2048, 1)).astype(np.float32), "customer_id": rng.integers(1, bench.V_CUSTOMERS + 1, size=(2048, 1)).astype(np.int32)}
pq = {k: torch.from_numpy(v).pin_memory() for k, v in h.items()}
for _ in range(5):
    index(pq)
def t(fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, qe = t(lambda: index._embed_queries(pq)); print(f"embed queries (H2D + tower): {ms:.3f} ms")
ms, (s, idx) = t(lambda: index.search(qe)); print(f"search: {ms:.3f} ms")
ms, _ = t(lambda: index._identifiers_dev[idx.long().clamp_(min=0)]); print(f"device id gather: {ms:.3f} ms")
ms, _ = t(lambda: idx.cpu().numpy()); print(f"idx.cpu().numpy(): {ms:.3f} ms")
ms, _ = t(lambda: index(pq)); print(f"index(pq) total: {ms:.3f} ms")
