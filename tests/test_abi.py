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


def test_header_is_plain_c(tmp_path):
    """include/hgin.h is the C ABI: it must compile as C99 on its own (no C++ / CUDA / torch types)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_header.c"
    src.write_text('#include "hgin.h"\nint main(void) { hgin_collate_field f; (void)f; return hgin_version() > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.dirname(_lib.HEADER_PATH),
                           "-fsyntax-only", str(src)])


def test_every_declared_symbol_is_exported(built):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.header_functions():
        assert hasattr(raw, name), name
    assert built.hgin_version() == 210


def test_argument_errors_do_not_need_a_gpu(built):
    # sizes are validated before any CUDA call: negative status + message, no exception, no abort
    assert built.hgin_csr_workspace_bytes(-1, 4) == -1
    rc = built.hgin_csr_build(None, 3, 0, 0, 1, 0, 0, None, None, None, None, None, 0, None)
    assert rc == -1 and b"index_bytes" in built.hgin_last_error()
    rc = built.hgin_gin_combine(4, None, None, -1, None, 0, 0, None, 0, 0, None, 0, 0, None, 0, None)
    assert rc == -1 and b"f_src" in built.hgin_last_error()
    with pytest.raises(_lib.HginError):
        _lib.check(rc, "hgin_gin_combine")


def test_argument_errors_of_the_newer_entry_points(built):
    """Validation happens before any CUDA call: status < 0 and a message, never an abort."""
    err = lambda: built.hgin_last_error().decode()
    rc = built.hgin_gin_combine_post(4, None, None, 0, None, 4, 4, None, 4, 4, None, 2, 0, None, 4, None, 4, 1, None, None, None, None, 0, None)
    assert rc == -1                                           # null x_self with SELF_CONCAT / post on a concat result
    rc = built.hgin_gin_combine_pre(4, None, None, 0, None, 4, 4, None, 4, 4, None, 0, 0, None, 4, 1, None, 0, None, None)
    assert rc == -1 and "slope" in err()                      # PReLU input activation without its slope
    rc = built.hgin_gin_combine_pre(4, None, None, 0, None, 4, 4, None, 4, 4, None, 0, 0, None, 4, 7, None, 0, None, None)
    assert rc == -1 and "input activation" in err()
    assert built.hgin_qt_baseline_workspace_bytes(-1, 0) == -1
    rc = built.hgin_qt_baseline(10, 4, 20, None, None, None, None, None, None, None, None, 0, None, None, None, 0, None)
    assert rc == -1 and "num_iterations" in err()
    rc = built.hgin_qt_baseline(10, 4, 20, None, None, None, None, None, None, None, None, 3, None, None, None, 0, None)
    assert rc == -1 and "null pointer" in err()
    rc = built.hgin_collate_offsets(4, None, 99, None, 10, None, None, None)
    assert rc == -1 and "bad sizes" in err()
    rc = built.hgin_collate_gather(70000, None, 10, 1, None, 1, None, 1, None)
    assert rc == -1 and "65535" in err()
    rc = built.hgin_linear_bwd_post(8, None, 4, None, 4, 0, None, None, 4, 4, None, 0, 0, None, 4, 0, 4, None, 4, None, None, None,
                                    None, 4, 9, None, None, None, 0, 0, None)
    assert rc == -1 and "post_act" in err()


def test_host_collate_runs_without_a_gpu_and_validates_ids(built):
    import ctypes as C
    import numpy as np
    ptr = np.array([0, 2, 5], dtype=np.int64)                 # two samples: 2 and 3 rows of width 2
    src = np.arange(10, dtype=np.int32)
    dst = np.zeros(10, dtype=np.int32)
    offsets = np.zeros(3, dtype=np.int64)
    field = (_lib.CollateField * 1)(_lib.CollateField(src.ctypes.data, dst.ctypes.data, ptr.ctypes.data, 2, 0, 0, 0))
    ids = np.array([1, 0], dtype=np.int32)
    rc = built.hgin_host_collate(2, ids.ctypes.data, 2, 1, field, 1, ptr.ctypes.data, offsets.ctypes.data, 2)
    assert rc == 0 and offsets.tolist() == [0, 3, 5]
    # sample 1 (rows 2..4) first, +0; then sample 0 (rows 0..1) with its row offset 3 added to every int32 entry
    assert dst.tolist() == [4, 5, 6, 7, 8, 9, 3, 4, 5, 6]
    bad = np.array([0, 2], dtype=np.int32)
    rc = built.hgin_host_collate(2, bad.ctypes.data, 2, 1, field, 1, ptr.ctypes.data, offsets.ctypes.data, 1)
    assert rc == -1 and b"outside" in built.hgin_last_error()


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
