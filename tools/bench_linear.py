"""Micro-benchmark of the dense-layer entry points for one shape (GPU): ms and effective GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_link_prediction_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2508800
k, n = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 128)
reps = 5
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(rows, k, device="cuda", generator=g)
W = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
b = torch.randn(n, device="cuda", generator=g)
a = torch.full((1,), 0.25, device="cuda")
gout = torch.randn(rows, n, device="cuda", generator=g)
dot_x = torch.randn(rows, k, device="cuda", generator=g)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


z, o = ops.linear_fwd(x, W, b, act=ops.ACT_PRELU, alpha=a, math_mode=ops.MATH_TF32)
t = timeit(lambda: ops.linear_fwd(x, W, b, act=ops.ACT_PRELU, alpha=a, math_mode=ops.MATH_TF32))
print(f"fwd tf32: {t:.3f} ms  {4 * rows * (k + 2 * n) / t / 1e6:.0f} GB/s")
for dot in (None, dot_x):
    t = timeit(lambda: ops.linear_bwd(gout, z, x, W, act=ops.ACT_PRELU, alpha=a, dot_x=dot, want_dalpha=True,
                                      math_mode=ops.MATH_TF32))
    units = 4 * rows * (2 * n + 2 * k + (k if dot is not None else 0))
    print(f"bwd tf32 (dot={'yes' if dot is not None else 'no'}): {t:.3f} ms  minimal-traffic {units / t / 1e6:.0f} GB/s")
