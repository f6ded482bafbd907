"""Micro-benchmark of tt_index_topk (development tool, GPU only): python scripts/index_microbench.py [n] [nq] [E] [K]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import torch
from pkg import _native as N

n = int(sys.argv[1]) if len(sys.argv) > 1 else 105542
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
E = int(sys.argv[3]) if len(sys.argv) > 3 else 64
K = int(sys.argv[4]) if len(sys.argv) > 4 else 100
lib = N.load()
g = torch.Generator(device="cuda").manual_seed(0)
C = torch.randn(n, E, device="cuda", generator=g).abs() * 0.1
Q = torch.relu(torch.randn(nq, E, device="cuda", generator=g) * 0.3)
rows_pad = ((n + 255) // 256 + 1) * 256
n_pad = 2 * rows_pad + rows_pad // 32 + 32
C32 = torch.empty_like(C); norms = torch.zeros(n_pad, device="cuda")
st = N.stream_ptr()
N.check(lib.tt_index_prepare(C.data_ptr(), E, n, E, C32.data_ptr(), norms.data_ptr(), st))
s = torch.empty(nq, K, device="cuda"); i = torch.empty(nq, K, dtype=torch.int32, device="cuda")
for impl, name in ((N.TT_IMPL_TC, "tensor-core filter"), (N.TT_IMPL_SIMT, "exact CUDA-core")):
    ws = torch.empty(int(lib.tt_index_workspace_bytes(nq, n, E, K, impl, 1)), dtype=torch.uint8, device="cuda")
    def run():
        N.check(lib.tt_index_topk(Q.data_ptr(), E, C.data_ptr(), E, C32.data_ptr(), norms.data_ptr(), nq, n, E, K, 0, s.data_ptr(), i.data_ptr(),
                                  ws.data_ptr(), ws.numel(), impl, st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10 if impl == N.TT_IMPL_TC else 3
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: n={n} nq={nq} E={E} K={K}: {ms:.3f} ms/batch, {nq / ms * 1e3:.0f} queries/s, {2.0 * nq * n * E / ms / 1e9:.1f} TFLOP/s algorithmic")

import ctypes
stages = (ctypes.c_float * 8)()
lib.tt_debug_index_stages(ctypes.cast(stages, ctypes.c_void_p))
ws = torch.empty(int(lib.tt_index_workspace_bytes(nq, n, E, K, N.TT_IMPL_TC, 1)), dtype=torch.uint8, device="cuda")
reps = 20
for _ in range(reps):
    N.check(lib.tt_index_topk(Q.data_ptr(), E, C.data_ptr(), E, C32.data_ptr(), norms.data_ptr(), nq, n, E, K, 0, s.data_ptr(), i.data_ptr(),
                              ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))
lib.tt_debug_index_stages(None)
names = ["prep", "filter", "select", "collect", "rescore", "fallback"]
print("stages (us, in situ): " + "  ".join(f"{nm} {stages[k] / reps * 1e3:.1f}" for k, nm in enumerate(names)))
