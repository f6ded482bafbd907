"""Development driver for the two-pass tcgen05 in-batch softmax (tt_tc_flash.cuh), GPU only.

    python scripts/flash_dev.py check [lbo sbo]     parity of tt_inbatch_softmax_step / _fwd / _bwd against a float64 torch evaluation
    python scripts/flash_dev.py time [B ...]        device time of the step at E = 64 (and 128), plus the in-kernel timeline of one CTA
The float64 evaluation here is a development aid; the parity tests proper (tests/test_gpu_tc.py) use the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hm-retrieval-two-tower_b200"))

import torch  # noqa: E402

from pkg import _native as N  # noqa: E402


def ref64(q, c, bias, off):
    q64, c64 = q.double(), c.double()
    z = q64 @ c64.T
    if bias is not None:
        z = z - bias.double()[None, :]
    lse = torch.logsumexp(z, dim=1)
    idx = torch.arange(q.shape[0], device=q.device)
    zd = z[idx, idx + off]
    loss = (lse - zd).sum()
    p = torch.exp(z - lse[:, None])
    p[idx, idx + off] -= 1.0
    return loss, lse, p @ c64, p.T @ q64


def run_step(lib, q, c, bias, off):
    Bq, E = q.shape
    Bc = c.shape[0]
    st = N.stream_ptr()
    lse = torch.zeros(Bq, device="cuda"); loss = torch.zeros(1, device="cuda")
    dq = torch.full((Bq, E), 7.0, device="cuda"); dc = torch.full((Bc, E), 7.0, device="cuda")
    ws = torch.empty(int(lib.tt_softmax_workspace_bytes(Bq, Bc, E)), dtype=torch.uint8, device="cuda")
    N.check(lib.tt_inbatch_softmax_step(q.data_ptr(), E, c.data_ptr(), E, bias.data_ptr() if bias is not None else None, Bq, Bc, E, off,
                                        lse.data_ptr(), loss.data_ptr(), dq.data_ptr(), E, dc.data_ptr(), E, ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))
    torch.cuda.synchronize()
    return loss, lse, dq, dc, ws


def rowerr(got, want):
    """largest row error relative to that row's norm (rows whose norm is above 1e-3 of the largest), and the overall max-abs error
    relative to the largest entry"""
    d = (got.double() - want).norm(dim=1)
    n = want.norm(dim=1)
    big = n > 1e-3 * n.max()
    return float((d[big] / n[big]).max()), float((got.double() - want).abs().max() / want.abs().max())


def check(lib, scale=0.3, trained=False):
    g = torch.Generator(device="cuda").manual_seed(0)
    ok = True
    shapes = [(256, 256, 64, 0), (128, 128, 64, 0), (1000, 1000, 64, 0), (130, 390, 64, 130), (3, 3, 64, 0), (2048, 2048, 64, 0),
              (300, 300, 128, 0), (257, 1000, 128, 5), (4096, 4096, 64, 0), (8192, 8192, 64, 0), (1024, 8192, 128, 1024)]
    for Bq, Bc, E, off in shapes:
        q = torch.relu(torch.randn(Bq, E, device="cuda", generator=g) * scale)
        c = torch.relu(torch.randn(Bc, E, device="cuda", generator=g) * scale)
        if trained:   # a sharp softmax: the positive dominates its row
            c[off:off + Bq] = q * 1.0 + 0.05 * c[off:off + Bq]
            q = q * 6.0
        bias = torch.log(torch.rand(Bc, device="cuda", generator=g) * 0.01 + 1e-5)
        loss, lse, dq, dc, _ = run_step(lib, q, c, bias, off)
        wl, wlse, wdq, wdc = ref64(q, c, bias, off)
        e_loss = abs(float(loss) - float(wl)) / abs(float(wl))
        e_lse = float((lse.double() - wlse).abs().max())
        eq, eqm = rowerr(dq, wdq)
        ec, ecm = rowerr(dc, wdc)
        good = e_loss < 1e-4 and e_lse < 1e-3 and eqm < 2e-3 and ecm < 2e-3
        ok &= good
        print(f"  {'ok ' if good else 'BAD'} Bq={Bq:5d} Bc={Bc:5d} E={E:3d} off={off:4d}: loss rel {e_loss:.2e}  lse abs {e_lse:.2e}  "
              f"dQ row-rel {eq:.2e} max-rel {eqm:.2e}  dC row-rel {ec:.2e} max-rel {ecm:.2e}", flush=True)
    return ok


def check_sep(lib):
    """fwd + bwd entry points (lse given) against the step"""
    g = torch.Generator(device="cuda").manual_seed(1)
    Bq, Bc, E, off = 700, 900, 64, 100
    q = torch.relu(torch.randn(Bq, E, device="cuda", generator=g) * 0.3)
    c = torch.relu(torch.randn(Bc, E, device="cuda", generator=g) * 0.3)
    bias = torch.log(torch.rand(Bc, device="cuda", generator=g) * 0.01 + 1e-5)
    loss, lse, dq, dc, ws = run_step(lib, q, c, bias, off)
    st = N.stream_ptr()
    lse2 = torch.zeros(Bq, device="cuda"); loss2 = torch.zeros(1, device="cuda")
    N.check(lib.tt_inbatch_softmax_fwd(q.data_ptr(), E, c.data_ptr(), E, bias.data_ptr(), Bq, Bc, E, off, lse2.data_ptr(), loss2.data_ptr(),
                                       ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))
    dq2 = torch.zeros(Bq, E, device="cuda"); dc2 = torch.zeros(Bc, E, device="cuda")
    N.check(lib.tt_inbatch_softmax_bwd(q.data_ptr(), E, c.data_ptr(), E, bias.data_ptr(), lse2.data_ptr(), Bq, Bc, E, off, dq2.data_ptr(), E,
                                       dc2.data_ptr(), E, ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))
    torch.cuda.synchronize()
    print(f"  sep vs step: lse equal {bool(torch.equal(lse, lse2))} loss {float(loss):.6f}/{float(loss2):.6f} "
          f"dQ max diff {float((dq - dq2).abs().max()):.3e} (scale {float(dq.abs().max()):.3e})  dC max diff {float((dc - dc2).abs().max()):.3e}")


def timeit(lib, B, E, n=20):
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.relu(torch.randn(B, E, device="cuda", generator=g) * 0.3)
    c = torch.relu(torch.randn(B, E, device="cuda", generator=g) * 0.3)
    bias = torch.log(torch.rand(B, device="cuda", generator=g) * 0.01 + 1e-5)
    st = N.stream_ptr()
    lse = torch.zeros(B, device="cuda"); loss = torch.zeros(1, device="cuda")
    dq = torch.empty(B, E, device="cuda"); dc = torch.empty(B, E, device="cuda")
    ws = torch.empty(int(lib.tt_softmax_workspace_bytes(B, B, E)), dtype=torch.uint8, device="cuda")

    def step():
        N.check(lib.tt_inbatch_softmax_step(q.data_ptr(), E, c.data_ptr(), E, bias.data_ptr(), B, B, E, 0, lse.data_ptr(), loss.data_ptr(),
                                            dq.data_ptr(), E, dc.data_ptr(), E, ws.data_ptr(), ws.numel(), N.TT_IMPL_TC, st))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 6.0 * B * B * E / (ms * 1e-3) / 1e12
    print(f"  step B={B} E={E}: {ms:.4f} ms  ({tf:.0f} TFLOP/s algorithmic, {tf / 1622:.3f} of the bf16 burst peak)", flush=True)
    return step


def timeline(lib, step, B, label):
    trace = torch.zeros(148 * 64 * 8, dtype=torch.int64, device="cuda")
    lib.tt_debug_flash(trace.data_ptr(), 0, 0)
    step()
    torch.cuda.synchronize()
    lib.tt_debug_flash(None, 0, 0)
    t = trace.cpu().numpy().reshape(148, 64, 8)
    # the trace buffer is shared by both passes of a step: the later launch (pass 2) overwrites pass 1
    import numpy as np
    first = np.array([t[c, 0, 0] if t[c, 0, 0] else 0 for c in range(148)], dtype=np.int64)
    last = t[:, :63, :].reshape(148, -1).max(axis=1)
    live = first > 0
    base = first[live].min()
    print(f"  {label}: CTAs with a trace {int(live.sum())}; first TMA issue after the earliest CTA (us): median {np.median(first[live] - base) / 1e3:.2f} "
          f"max {(first[live] - base).max() / 1e3:.2f}; last stamp: median {np.median(last[live] - base) / 1e3:.2f} max {(last[live] - base).max() / 1e3:.2f}; "
          f"units per CTA {int((t[:, :, 3] > 0).sum(axis=1)[live].min())}..{int((t[:, :, 3] > 0).sum(axis=1)[live].max())}")
    slow = int(np.argmax(np.where(live, last, 0)))
    ent, ext = t[:, 63, 0][live], t[:, 63, 1][live]
    print(f"  {label}: CTA entry after the earliest entry (us): median {np.median(ent - ent.min()) / 1e3:.2f} max {(ent - ent.min()).max() / 1e3:.2f}; "
          f"first TMA issue after own entry: median {np.median(first[live] - ent) / 1e3:.2f}; exit after the earliest entry: median {np.median(ext - ent.min()) / 1e3:.2f} "
          f"max {(ext - ent.min()).max() / 1e3:.2f}")
    t[:, 63, :] = 0
    # long gaps of warp 0 between finishing one unit and seeing the next S, or inside a unit: (cta, unit, kind, gap us, at us)
    ev = []
    for cta in range(148):
        for u in range(1, 16):
            if t[cta, u, 2] and t[cta, u - 1, 3] and t[cta, u, 2] - t[cta, u - 1, 3] > 1000:
                ev.append((cta, u, 'wait', (t[cta, u, 2] - t[cta, u - 1, 3]) / 1e3, (t[cta, u, 2] - base) / 1e3))
        for u in range(0, 16):
            if t[cta, u, 2] and t[cta, u, 3] and t[cta, u, 3] - t[cta, u, 2] > 2200:
                ev.append((cta, u, 'epi', (t[cta, u, 3] - t[cta, u, 2]) / 1e3, (t[cta, u, 3] - base) / 1e3))
    print(f"  {label}: {len(ev)} long gaps of warp 0:", ' '.join(f'({c},{u},{k},{g:.1f}us@{a:.1f})' for c, u, k, g, a in ev[:60]))
    for cta in (0, 73, slow):
        tt = t[cta]
        n = int((tt[:, 3] > 0).sum())
        if n == 0:
            continue
        t0 = tt[0, 0]
        print(f"  {label} cta {cta}: {n} traced units; us since the first TMA issue: unit: tma | mma1(s0,sub0) issued, complete | s_seen(w0) | s_done(w0) | mma2(s0,sub1) issued, complete")
        for u in list(range(min(n, 16))):
            print(f"    {u:3d}: " + " ".join(f"{(tt[u, e] - t0) / 1e3:7.2f}" if tt[u, e] else "      -" for e in (0, 1, 4, 2, 3, 7, 5)))


def main():
    lib = N.load()
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        if len(sys.argv) > 3:
            lib.tt_debug_flash(None, int(sys.argv[2]), int(sys.argv[3]))
        print("random ReLU operands:")
        ok = check(lib)
        print("sharp softmax (positive dominates):")
        ok &= check(lib, trained=True)
        check_sep(lib)
        print("ALL OK" if ok else "FAILED")
        sys.exit(0 if ok else 1)
    if mode == "slow":   # every chunk through the checked path (cost of the out-of-line code when it is warm)
        lib.tt_debug_flash(None, 0, -1)
        step = timeit(lib, 8192, 64, n=5)
        trace = torch.zeros(148 * 64 * 8, dtype=torch.int64, device="cuda")
        lib.tt_debug_flash(trace.data_ptr(), 0, -1)
        step()
        torch.cuda.synchronize()
        lib.tt_debug_flash(None, 0, 0)
        t = trace.cpu().numpy().reshape(148, 64, 8)
        for cta in (5, 100):
            tt = t[cta]
            print(f"  cta {cta} (every chunk checked): unit: s_seen -> s_done (us)")
            print("    " + " ".join(f"{(tt[u, 3] - tt[u, 2]) / 1e3:5.2f}" for u in range(14) if tt[u, 3]))
        return
    if mode == "inorder":   # first product issued right behind the second one (no completion wait): parity, then time
        lib.tt_debug_flash(None, -1, 0)
        ok = check(lib)
        print("in-order issue: parity", "ok" if ok else "FAILED")
        for B in (8192, 65536):
            timeit(lib, B, 64, n=20 if B <= 16384 else 5)
        lib.tt_debug_flash(None, 0, 0)
        return
    if mode == "once":   # a few steps only (under ncu)
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
        timeit(lib, B, 64, n=2)
        return
    sizes = [int(a) for a in sys.argv[2:]] or [8192, 16384, 65536]
    for B in sizes:
        step = timeit(lib, B, 64, n=20 if B <= 16384 else 5)
        if B == sizes[0]:
            timeline(lib, step, B, f"B={B}")
    timeit(lib, 8192, 128)


if __name__ == "__main__":
    main()
