"""libtt.so builds here (no GPU), loads, and exports every symbol include/tt.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from pkg import _native

    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(_native.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in tt.h but not exported: {missing}"
    assert sorted(_native.SIGNATURES) == names, "pkg/_native.py binds a different set than tt.h declares"


def test_version_and_error_string(lib):
    assert lib.tt_version() >= 100
    assert isinstance(lib.tt_last_error(), bytes)


def test_struct_layout_matches_header(lib):
    from pkg import _native as N

    assert ctypes.sizeof(N.TTFeature) == 32
    assert ctypes.sizeof(N.TTSparseJob) == 24 + 24 + 8 * N.TT_MAX_SRC * 2 + 4 * N.TT_MAX_SRC


def test_argument_errors_do_not_need_a_gpu(lib):
    # pure argument validation happens before any CUDA call
    rc = lib.tt_dense_adagrad(None, None, None, 10, 0.1, 1e-7, None)
    assert rc == -1 and b"null pointer" in lib.tt_last_error()
    rc = lib.tt_recall_hits(None, 1, None, 1, None, 1, None, None)
    assert rc == -1


def test_sparse_plan_access_width_follows_alignment(lib):
    """Host half of the vectorised segmented reduce (csrc/tt_sparse.cu make_plans): a 64 / 128-bit access is only planned when the
    width divides e and every base pointer the kernel offsets by whole rows is aligned for it.  No CUDA call, fabricated addresses."""
    from pkg import _native as N

    A = 1 << 20                                     # a 1 MiB-aligned fake device address

    def plan(e, table=A, slot0=2 * A, slot1=0, grad=3 * A, grad_ld=None, ws=8 * A, shard_world=0, nsrc=1):
        jobs = (N.TTSparseJob * 1)()
        j = jobs[0]
        j.table, j.slot0, j.slot1 = table, slot0, slot1 or None
        j.rows, j.e, j.nsrc, j.n_per_src, j.shard_rank, j.shard_world = 1000, e, nsrc, 256, 0, shard_world
        for s in range(nsrc):
            j.ids[s], j.grad[s], j.grad_ld[s] = 4 * A, grad + 4 * e * s, grad_ld or e * nsrc
        nbytes = lib.tt_sparse_workspace_bytes(1, 256 * nsrc, e) + 64
        vec, gvec = (ctypes.c_int32 * 1)(), (ctypes.c_int32 * 1)()
        rc = lib.tt_debug_sparse_plan(jobs, 1, ws, nbytes, vec, gvec)
        assert rc == 0, lib.tt_last_error()
        return vec[0], gvec[0]

    assert plan(8) == (1, 1) and plan(16) == (1, 1) and plan(32) == (1, 1)
    assert plan(64) == (2, 1) and plan(48) == (2, 1)
    assert plan(128) == (4, 1) and plan(96) == (4, 1) and plan(256) == (4, 1)
    assert plan(33) == (1, 1) and plan(66) == (2, 1) and plan(130) == (2, 1)          # narrowed until the width divides e
    assert plan(64, grad=3 * A + 4) == (2, 0)                                         # dX slice starting at an odd column
    assert plan(64, grad_ld=65) == (2, 0) and plan(128, grad_ld=130) == (4, 0) and plan(128, grad=3 * A + 8) == (4, 0)
    assert plan(64, nsrc=2) == (2, 1)                                                  # two sources side by side in one dX block
    assert plan(128, table=A + 8) == (2, 1) and plan(128, slot0=2 * A + 4) == (1, 1) and plan(128, slot1=5 * A + 8) == (2, 1)
    assert plan(64, ws=8 * A + 4)[0] == 1                                              # piece buffers live in the workspace
    assert plan(64, grad=3 * A + 4, shard_world=2) == (2, 1)                           # sharded: rows are read from the staged copy
    vec = (ctypes.c_int32 * 1)()
    assert lib.tt_debug_sparse_plan(None, 1, A, 1 << 20, vec, vec) == -1
