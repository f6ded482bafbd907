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
