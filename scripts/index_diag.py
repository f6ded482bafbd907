"""Diagnostic (GPU): score distribution of the bench index scenario vs the TF32 filter window."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import numpy as np, torch
import bench

model = bench.build_gpu_model()
rng = np.random.default_rng(0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    model.train_step({k: torch.from_numpy(v).cuda() for k, v in bench.make_batch(rng, 8192).items()})
art = np.arange(1, bench.V_ARTICLES + 1, dtype=np.int32)
embs = []
for lo in range(0, bench.V_ARTICLES, 10000):
    a = art[lo:lo + 10000]
    embs.append(model.candidate_tower({"article_id": a.reshape(-1, 1), "product_type_name": (a % 131 + 1).reshape(-1, 1), "colour_group_name": (a % 50 + 1).reshape(-1, 1)}))
C = torch.cat(embs)
q = model.query_tower({"age": rng.random((256, 1)).astype(np.float32), "customer_id": rng.integers(1, bench.V_CUSTOMERS + 1, size=(256, 1)).astype(np.int32)})
S = q @ C.T
cn = C.norm(dim=1); qn = q.norm(dim=1)
print("corpus norms: min %.4g median %.4g max %.4g; nonzero rows %d / %d" % (cn.min(), cn.median(), cn.max(), int((cn > 0).sum()), C.shape[0]))
print("distinct corpus rows:", torch.unique(C, dim=0).shape[0])
print("query norms: min %.4g median %.4g max %.4g" % (qn.min(), qn.median(), qn.max()))
K = 100
top, _ = S.topk(K, dim=1)
sK = top[:, -1:]
eps = (2.0 ** -9) * qn[:, None] * cn[None, :]
within = (S + 2 * eps >= sK).sum(dim=1)
print("s_K median %.5g; candidates with s + 2eps >= s_K: min %d median %d max %d" % (sK.median(), within.min(), within.median(), within.max()))
rel = (S >= sK * (1 - 2.0 ** -8)).sum(dim=1)
print("candidates within 0.4%% of s_K: median %d" % rel.median())
print("exact ties with s_K: median %d" % (S == sK).sum(dim=1).median())

# ---- internals of the tensor-core filter on the same data -------------------------------------------
from pkg import _native as N
from pkg.modelling.indices.brute_force import BruteForceIndex
lib = N.load()
nq = 2048
qx = {"age": rng.random((nq, 1)).astype(np.float32), "customer_id": rng.integers(1, bench.V_CUSTOMERS + 1, size=(nq, 1)).astype(np.int32)}
index = BruteForceIndex(100, model.query_tower, [(art, C)])
qe = index._embed_queries(qx)
s, i = index.search(qe)
torch.cuda.synchronize()
E, n, K = 64, C.shape[0], 100
def al(x): return (x + 255) // 256 * 256
n_tiles = (n + 255) // 256; ngroups = n_tiles * 8; cap = 4 * K + 512
off = 0
offs = {}
for name, size in (("q32", nq * E * 4), ("eps", nq * 4), ("thr", nq * 4), ("gmax", nq * ngroups * 4), ("cnt", nq * 4), ("flags", nq * 4), ("cand", nq * cap * 4)):
    offs[name] = off; off += al(size)
ws = index._ws
def view(name, dtype, count): return ws[offs[name]:offs[name] + count * 4].view(dtype)
thr = view("thr", torch.float32, nq); cnt = view("cnt", torch.int32, nq); flags = view("flags", torch.int32, nq); kap = view("eps", torch.float32, nq)
gmax = view("gmax", torch.float32, nq * ngroups).view(nq, ngroups)
print("cnt: min %d median %d max %d; flagged %d / %d" % (cnt.min(), cnt.median(), cnt.max(), int(flags.sum()), nq))
S2 = qe @ C.T
cn2 = C.norm(dim=1)
Lb = S2 - kap[:, None] * cn2[None, :] * 1.0001
real_groups = (n + 31) // 32
pad = real_groups * 32 - n
Lp = torch.cat([Lb, torch.full((nq, pad), float("-inf"), device="cuda")], dim=1).view(nq, real_groups, 32).max(dim=2).values
lam = Lp.topk(K, dim=1).values[:, -1]
print("lambda (torch fp32) vs thr (kernel): max abs diff %.4g; thr min %.4g median %.4g; lam median %.4g" % ((lam - thr).abs().max(), thr.min(), thr.median(), lam.median()))
print("gmax vs torch group max: max abs diff %.4g" % (gmax[:, :real_groups] - Lp).abs().max())
print("kappa median %.4g (expected %.4g)" % (kap.median(), (2.0 ** -9) * qe.norm(dim=1).median()))
