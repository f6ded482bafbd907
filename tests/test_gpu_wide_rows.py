"""Sparse Adagrad on embedding rows wider than 64 columns (BASELINE configs[4]: dim 128) against the oracle, bit for bit.
Rows of 65..256 columns take 16-byte accesses in the segmented reduce (one or two per lane); widths the access does not divide
take the scalar body; a feature's gradient slice may start 16-, 8- or 4-byte aligned inside dX (vector or scalar gradient loads).
Kept in its own file, after the other GPU suites in collection order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import two_tower_oracle as O  # noqa: E402


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("goff", [4, 2, 1])
def test_sparse_adagrad_wide_rows_bit_exact(lib, goff):
    import torch

    from pkg import _native as N

    rng = np.random.default_rng(70 + goff)
    B = 3000
    specs = [(5000, 128, 1), (700, 256, 1), (4000, 96, 2), (900, 130, 1), (20_000, 64, 1)]     # (rows, e, gradient sources)
    jobs = (N.TTSparseJob * len(specs))()
    keep, host = [], []
    for j, (rows, e, nsrc) in enumerate(specs):
        table = rng.standard_normal((rows, e)).astype(np.float32)
        acc = np.full((rows, e), 0.1, np.float32)
        ld = e + 4                                                      # the slice sits at column goff of a (B, e + 4) block
        ids = [np.minimum(rng.zipf(1.3, size=B) - 1, rows - 1).astype(np.int32) for _ in range(nsrc)]   # heavy duplicates
        for i in ids:
            i[3] = rows + 5                                             # out of range: folds into the OOV row 0
        grads = [rng.standard_normal((B, ld)).astype(np.float32) for _ in range(nsrc)]
        dt, da = _dev(torch, table), _dev(torch, acc)
        dids, dgs = [_dev(torch, i) for i in ids], [_dev(torch, g) for g in grads]
        keep += [dt, da] + dids + dgs
        jobs[j].table, jobs[j].slot0, jobs[j].slot1 = dt.data_ptr(), da.data_ptr(), None
        jobs[j].rows, jobs[j].e, jobs[j].nsrc, jobs[j].n_per_src = rows, e, nsrc, B
        for s in range(nsrc):
            jobs[j].ids[s], jobs[j].grad[s], jobs[j].grad_ld[s] = dids[s].data_ptr(), dgs[s].data_ptr() + 4 * goff, ld
        ids_all = np.concatenate([np.where((i < 0) | (i >= rows), 0, i) for i in ids])
        vals_all = np.concatenate([g[:, goff:goff + e] for g in grads], axis=0)
        host.append((table, acc, O.IndexedSlices(ids_all, np.ascontiguousarray(vals_all)), dt, da))
    ws = torch.empty(int(lib.tt_sparse_workspace_bytes(len(specs), 2 * B, 256)), dtype=torch.uint8, device="cuda")
    st = N.stream_ptr()
    N.check(lib.tt_sparse_sort(jobs, len(specs), ws.data_ptr(), ws.numel(), st), "tt_sparse_sort")
    N.check(lib.tt_sparse_adagrad(jobs, len(specs), 0.05, 1e-7, ws.data_ptr(), ws.numel(), st), "tt_sparse_adagrad")
    torch.cuda.synchronize()
    for table, acc, slices, dt, da in host:
        O.adagrad_sparse(table, acc, slices, 0.05)
        assert np.array_equal(dt.cpu().numpy(), table)                  # same summation order, same rounding
        assert np.array_equal(da.cpu().numpy(), acc)
