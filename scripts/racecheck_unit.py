#!/usr/bin/env python
"""Unit-shape invocations of every hot-path entry point, run under `compute-sanitizer --tool racecheck` (shared-memory hazards)
and `--tool memcheck`:   compute-sanitizer --tool racecheck python scripts/racecheck_unit.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import torch  # noqa: E402

from pkg import _native as N  # noqa: E402
from pkg.modelling._device import set_seed  # noqa: E402
from pkg.modelling.indices.brute_force import BruteForceIndex  # noqa: E402
from pkg.modelling.models.two_tower_model import TwoTowerModel  # noqa: E402
from pkg.modelling.optimizer_factory import OptimizerFactory  # noqa: E402
from pkg.schema import dtypes as tt  # noqa: E402
from pkg.schema.features import Feature, FeatureFamily  # noqa: E402

lib = N.load()
set_seed(3)
vq, vc, B = 500, 300, 256
qf = [Feature("age", tt.float32, FeatureFamily.QUERY), Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=64)]
cf = [Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=64), Feature("colour_group_name", tt.string, FeatureFamily.CANDIDATE, embedding_size=8)]
qf[1].set_vocab_size(vq); cf[0].set_vocab_size(vc); cf[1].set_vocab_size(50)
for impl in (N.TT_IMPL_AUTO, N.TT_IMPL_SIMT):
    m = TwoTowerModel(qf, cf, "article_id", 64, [96], [96], candidate_prob_lookup={str(i + 1): 1.0 / vc for i in range(vc)})
    m.impl = impl
    m.use_cuda_graph = False
    m.compile(optimizer=OptimizerFactory.get_optimizer("adagrad", {"learning_rate": 0.05}))
    rng = np.random.default_rng(0)
    for _ in range(2):
        art = rng.integers(1, vc + 1, size=(B, 1)).astype(np.int32)
        loss = float(m.train_step({"age": rng.random((B, 1)).astype(np.float32), "customer_id": rng.integers(0, vq + 1, size=(B, 1)).astype(np.int32),
                                   "article_id": art, "colour_group_name": (art % 50 + 1).astype(np.int32)})["loss"])
    print("train_step impl", impl, "loss", loss, flush=True)
    art = np.arange(1, vc + 1, dtype=np.int32).reshape(-1, 1)
    emb = m.candidate_tower({"article_id": art, "colour_group_name": (art % 50 + 1).astype(np.int32)})
    index = BruteForceIndex(10, m.query_tower, [(art.reshape(-1), emb)])
    index.impl = impl
    ids = index({"age": rng.random((40, 1)).astype(np.float32), "customer_id": rng.integers(0, vq + 1, size=(40, 1)).astype(np.int32)})
    print("index impl", impl, ids[0, :3], flush=True)
# a corpus large enough for the tensor-core filter path
g = torch.Generator(device="cuda").manual_seed(1)
corpus = torch.randn((4096, 64), generator=g, device="cuda") * 0.25


class Ident:
    def __call__(self, x):
        return x["q"]

    def get_input_signature(self):
        return {}


big = BruteForceIndex.from_local_rows(16, Ident(), corpus, 0, 4096)
s, i = big.search(torch.randn((130, 64), generator=g, device="cuda"))
torch.cuda.synchronize()
print("filter index", i[0, :3].tolist(), flush=True)
