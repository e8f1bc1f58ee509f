"""Diagnostic (GPU): which UMMA shared-memory descriptor fields make the MN-major tf32
weight-gradient kernel (gemm_tn_kernel) agree with a^T b?  Sweeps TMA swizzle mode x UMMA layout
type x SBO x LBO x K-step and prints the max abs error of each combination."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_link_prediction_b200 import ops  # noqa: E402

torch.manual_seed(0)
rows, n, k = 4096, 128, 128
a = torch.randn(rows, n, device="cuda")
b = torch.randn(rows, k, device="cuda")
ref = (a.double().t() @ b.double()).float()
scale = float(ref.abs().max())
results = []
# CUtensorMapSwizzle: 3 = 128B (16B chunks), 4 = 128B_ATOM_32B;  UMMA layout: 2 = SW128, 1 = SW128_BASE32B
for swz, lt in [(4, 1), (3, 2), (4, 2), (3, 1)]:
    for sbo, lbo, kstep in itertools.product([512, 1024, 256], [4096, 1024, 128], [1024, 512]):
        try:
            out = ops.debug_gemm_tn(a, b, tma_swizzle=swz, lbo=lbo, sbo=sbo, layout_type=lt, k_step_bytes=kstep)
            torch.cuda.synchronize()
            err = float((out - ref).abs().max()) / scale
        except Exception as e:  # noqa: BLE001
            err = float("nan")
            print("exception", swz, lt, sbo, lbo, kstep, e)
        results.append((err, swz, lt, sbo, lbo, kstep))
results.sort(key=lambda r: (r[0] != r[0], r[0]))
print("rel_err  tma_swizzle layout_type sbo lbo k_step")
for r in results[:12]:
    print("%.3e  %d %d %d %d %d" % r)
print("default :", float((ops.debug_gemm_tn(a, b) - ref).abs().max()) / scale)
