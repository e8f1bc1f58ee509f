"""TEST INFRASTRUCTURE: ctypes binding of oracle/hgin_oracle.c (numpy in, numpy out)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhgin_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "hgin_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libhgin_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_csr_build.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def csr_build(edge_index, num_src, num_dst):
    """edge_index: int64 [2,E] numpy.  Returns (rowptr i32[num_dst+1], col i32[E], perm i32[E])."""
    ei = np.ascontiguousarray(edge_index, dtype=np.int64)
    E = ei.shape[1]
    rowptr = np.zeros(num_dst + 1, np.int32)
    col = np.zeros(E, np.int32)
    perm = np.zeros(E, np.int32)
    rc = lib().oracle_csr_build(ctypes.c_int64(E), _p(ei[0]), _p(ei[1]), ctypes.c_int64(num_src),
                                ctypes.c_int64(num_dst), _p(rowptr), _p(col), _p(perm))
    if rc != 0:
        raise ValueError(f"oracle_csr_build failed ({rc}): index out of range")
    return rowptr, col, perm


def gin_combine(rowptr, col, x_src, x_dst, eps, concat):
    """x_src f32 [Ns,Fs] (any row stride), x_dst f32 [Nd,Fd] or None.  Returns h f32."""
    assert x_src.dtype == np.float32 and x_src.strides[1] == 4
    nd = rowptr.shape[0] - 1
    fs = x_src.shape[1]
    fd = 0 if x_dst is None else x_dst.shape[1]
    width = fs + fd if (concat and x_dst is not None) else fs
    if x_dst is not None and not concat:
        assert fd == fs
    h = np.zeros((nd, width), np.float32)
    lib().oracle_gin_combine(
        ctypes.c_int64(nd), _p(rowptr), _p(col), _p(x_src), ctypes.c_int64(x_src.strides[0] // 4),
        ctypes.c_int32(fs), _p(x_dst) if x_dst is not None else None,
        ctypes.c_int64(0 if x_dst is None else x_dst.strides[0] // 4), ctypes.c_int32(fd),
        ctypes.c_float(eps), ctypes.c_int32(1 if concat else 0), _p(h), ctypes.c_int64(width))
    return h


def gather_t(rowptr_t, col_t, g):
    assert g.dtype == np.float32 and g.strides[1] == 4
    ns = rowptr_t.shape[0] - 1
    dx = np.zeros((ns, g.shape[1]), np.float32)
    lib().oracle_gather_t(ctypes.c_int64(ns), _p(rowptr_t), _p(col_t), _p(g),
                          ctypes.c_int64(g.strides[0] // 4), ctypes.c_int32(g.shape[1]), _p(dx),
                          ctypes.c_int64(g.shape[1]))
    return dx
