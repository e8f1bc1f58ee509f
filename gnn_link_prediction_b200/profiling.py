"""Per-kernel CUDA-event timing used by bench.py to attribute step time and to compute the
roofline of the aggregation kernel live (events are recorded on the launching stream)."""
from __future__ import annotations

import torch

L2_BYTES = 126 * 1024 * 1024  # B200 L2


class KernelTimer:
    """`with timer.region(name, **meta)` brackets one C-ABI call with two CUDA events."""

    def __init__(self):
        self.records = []   # (name, start, stop, meta)
        self.enabled = True

    def region(self, name, **meta):
        return _Region(self, name, meta)

    def summary(self):
        """{name: dict(ms=total, launches=count, **summed numeric meta)} after a device sync."""
        torch.cuda.synchronize()
        out = {}
        for name, a, b, meta in self.records:
            d = out.setdefault(name, {"ms": 0.0, "launches": 0})
            d["ms"] += a.elapsed_time(b)
            d["launches"] += 1
            for k, v in meta.items():
                d[k] = d.get(k, 0) + v
        return out

    def clear(self):
        self.records.clear()


class _Region:
    def __init__(self, timer, name, meta):
        self.t, self.name, self.meta = timer, name, meta

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        self.b.record()
        self.t.records.append((self.name, self.a, self.b, self.meta))


def combine_bytes(num_rows, num_cols, num_edges, f_src, f_self, width, elem_bytes=4):
    """Algorithmic and compulsory bytes of one hgin_gin_combine launch, rows stored in
    `elem_bytes` (4 = fp32, 2 = bf16), indices int32:
    B_alg = E*(F*s + 4) + (rows+1)*4 + rows*F_self*s + rows*F_out*s; the compulsory variant
    charges each source row once (min(E, N_src)) instead of once per edge."""
    s = elem_bytes
    fixed = num_edges * 4 + (num_rows + 1) * 4 + num_rows * f_self * s + num_rows * width * s
    alg = num_edges * f_src * s + fixed
    compulsory = min(num_edges, num_cols) * f_src * s + fixed
    return alg, compulsory
