#!/usr/bin/env python
"""Development timing of BruteForceIndex.search (2048 queries, K=100) at a few corpus sizes; prints ms per batch and the phase split.
    python scripts/index_time.py 105542 10000000"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import torch  # noqa: E402

from pkg.modelling.indices.brute_force import BruteForceIndex  # noqa: E402



class Ident:
    def __call__(self, x):
        return x["q"]

    def get_input_signature(self):
        return {}


sizes = [int(float(a)) for a in sys.argv[1:]] or [105542]
g = torch.Generator(device="cuda").manual_seed(3)
q = torch.randn((2048, 64), generator=g, device="cuda").abs() * 0.3
for n in sizes:
    corpus = torch.randn((n, 64), generator=g, device="cuda") * 0.25
    index = BruteForceIndex.from_local_rows(100, Ident(), corpus, 0, n)
    for _ in range(3):
        index.search(q)
    torch.cuda.synchronize()
    reps = 50 if n < 1e6 else 10
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        index.search(q)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"n={n} fine={os.environ.get('TT_IDX_FINE', '-')} sample_min={os.environ.get('TT_IDX_SAMPLE_MIN', '-')}: {ms:.4f} ms/batch = {2048 / ms * 1e3 / 1e6:.2f} M q/s", flush=True)
    del index, corpus
