"""The C-ABI library loads and exports exactly what include/hgin.h declares (no compute calls)."""
import ctypes
import os

import pytest

from gnn_link_prediction_b200 import _lib


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert sorted(_lib.SIGNATURES) == _lib.header_functions()


def test_every_declared_symbol_is_exported(built):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.header_functions():
        assert hasattr(raw, name), name
    assert built.hgin_version() == 101


def test_argument_errors_do_not_need_a_gpu(built):
    # sizes are validated before any CUDA call: negative status + message, no exception, no abort
    assert built.hgin_csr_workspace_bytes(-1, 4) == -1
    rc = built.hgin_csr_build(None, 3, 0, 0, 1, 0, 0, None, None, None, None, None, 0, None)
    assert rc == -1 and b"index_bytes" in built.hgin_last_error()
    rc = built.hgin_gin_combine(4, None, None, -1, None, 0, 0, None, 0, 0, None, 0, 0, None, 0, None)
    assert rc == -1 and b"f_src" in built.hgin_last_error()
    with pytest.raises(_lib.HginError):
        _lib.check(rc, "hgin_gin_combine")


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhgin.so")
    with pytest.raises(_lib.HginError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected(built):
    import torch
    from gnn_link_prediction_b200 import ops
    with pytest.raises(_lib.HginError, match="no CPU fallback"):
        ops.csr_build(torch.zeros(2, 3, dtype=torch.int64), 4, 4)
    with pytest.raises(_lib.HginError, match="no CPU fallback"):
        ops.linear_fwd(torch.zeros(4, 3), torch.zeros(2, 3))
