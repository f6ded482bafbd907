#!/usr/bin/env python
"""Development check of the two-pass softmax at the cross-GPU shapes (B rows x G.B columns, diagonal at rank.B) against float64,
with near-initialisation operands (small tower outputs, near-uniform softmax) and with the element-wise error that Adagrad sees."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))
import torch  # noqa: E402

from flash_dev import ref64, run_step  # noqa: E402
from pkg import _native as N  # noqa: E402

lib = N.load()
g = torch.Generator(device="cuda").manual_seed(0)
for scale in (0.3, 0.02):
    for (Bq, Bc, E) in ((2048, 16384, 64), (2048, 4096, 64), (16384, 16384, 64), (2048, 16384, 128)):
        for off in sorted({0, Bc - Bq, (Bc - Bq) // 2 // 2048 * 2048}):
            q = torch.relu(torch.randn(Bq, E, device="cuda", generator=g) * scale)
            c = torch.relu(torch.randn(Bc, E, device="cuda", generator=g) * scale)
            bias = torch.log(torch.rand(Bc, device="cuda", generator=g) * 0.01 + 1e-5)
            loss, lse, dq, dc, _ = run_step(lib, q, c, bias, off)
            wl, wlse, wdq, wdc = ref64(q, c, bias, off)
            eq = float((dq.double() - wdq).abs().max()); ec = float((dc.double() - wdc).abs().max())
            print(f"scale {scale} Bq={Bq} Bc={Bc} E={E} off={off}: loss rel {abs(float(loss) - float(wl)) / abs(float(wl)):.1e}  "
                  f"dQ max abs err {eq:.2e} (max |dQ| {float(wdq.abs().max()):.2e})  dC max abs err {ec:.2e} (max |dC| {float(wdc.abs().max()):.2e})", flush=True)
