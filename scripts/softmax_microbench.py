"""Micro-benchmark + in-kernel timeline of the tcgen05 in-batch softmax (development tool, GPU only).

    python scripts/softmax_microbench.py [B] [E]
Prints ms for fwd / bwd at several split caps and, for one configuration, the per-CTA timeline recorded by
the kernel (globaltimer stamps; see RowPanelParams::trace)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))

import torch  # noqa: E402

from pkg import _native as N  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    lib = N.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.relu(torch.randn(B, E, device="cuda", generator=g) * 0.3)
    c = torch.relu(torch.randn(B, E, device="cuda", generator=g) * 0.3)
    q32, c32 = torch.empty_like(q), torch.empty_like(c)
    st = N.stream_ptr()
    N.check(lib.tt_round_tf32(q.data_ptr(), E, q32.data_ptr(), E, B, E, st))
    N.check(lib.tt_round_tf32(c.data_ptr(), E, c32.data_ptr(), E, B, E, st))
    bias = torch.log(torch.rand(B, device="cuda", generator=g) * 0.01 + 1e-5)
    lse = torch.empty(B, device="cuda"); loss = torch.zeros(1, device="cuda")
    dq = torch.empty(B, E, device="cuda"); dc = torch.empty(B, E, device="cuda")
    ws = torch.empty(int(lib.tt_softmax_workspace_bytes(B, B, E)), dtype=torch.uint8, device="cuda")

    def fwd():
        N.check(lib.tt_inbatch_softmax_fwd(q32.data_ptr(), E, c32.data_ptr(), E, bias.data_ptr(), B, B, E, 0, lse.data_ptr(), loss.data_ptr(),
                                           ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))

    def bwd():
        N.check(lib.tt_inbatch_softmax_bwd(q32.data_ptr(), E, c32.data_ptr(), E, bias.data_ptr(), lse.data_ptr(), B, B, E, 0, dq.data_ptr(), E,
                                           dc.data_ptr(), E, ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))

    def time_ms(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    print(f"B={B} E={E}")
    for cap in (1, 2, 4, 8, 16):
        lib.tt_debug_tc(None, cap)
        print(f"  max_splits={cap:2d}: fwd {time_ms(fwd):8.3f} ms   bwd(2 passes) {time_ms(bwd):8.3f} ms")
    # stream-K timeline: [cta][64 units][8] (see SkParams::trace)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        trace = torch.zeros(148 * 64 * 8, dtype=torch.int64, device="cuda")
        lib.tt_debug_tc(trace.data_ptr(), 0)
        fn()
        torch.cuda.synchronize()
        lib.tt_debug_tc(None, 0)
        t = trace.cpu().numpy().reshape(148, 64, 8)
        for cta in (0, 73):
            tt = t[cta]
            n = int((tt[:, 3] > 0).sum())
            if n == 0:
                continue
            t0 = tt[0, 0]
            print(f"{name} cta {cta}: {n} traced units; us since the first TMA issue: unit: tma mma1 s_seen s_released | mma warp: t_landed s_free p_ready mma2")
            for u in list(range(min(n, 14))) + ([n - 1] if n > 14 else []):
                print(f"    {u:3d}: " + " ".join(f"{(tt[u, e] - t0) / 1e3:7.2f}" if tt[u, e] else "      -" for e in range(8)))


if __name__ == "__main__":
    main()
