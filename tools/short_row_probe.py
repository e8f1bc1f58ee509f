"""Probe of the short-row aggregation kernels on a Cfg-C shaped link->path relation: per-launch time and achieved
bandwidth over the compulsory bytes, in the forward (self rows) and backward (self + post-activation rows) modes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_link_prediction_b200 import ops                      # noqa: E402
from gnn_link_prediction_b200.functional import GraphCSR      # noqa: E402
from tools.staged_probe import relation, timed                # noqa: E402


def main():
    blocks, n_path, n_link, hops = 1024, 2450, 200, 3
    ei, ptr_p, ptr_l = relation(blocks, n_path, n_link, hops, 0)      # path -> link edges; the short side is its transpose
    et = ("path", "to", "link")
    graph = GraphCSR({et: ei.cuda()}, {"path": blocks * n_path, "link": blocks * n_link})
    csr = graph.bwd(et)               # rows = path, ~3 link neighbours each
    for dt in (torch.float32, torch.bfloat16):
        es = 4 if dt == torch.float32 else 2
        x = torch.randn(blocks * n_link, 128, device="cuda").to(dt)
        xs = torch.randn(blocks * n_path, 128, device="cuda").to(dt)
        z = torch.randn(blocks * n_path, 128, device="cuda").to(dt)
        out = torch.empty(blocks * n_path, 128, device="cuda", dtype=dt)
        eps = torch.tensor([0.1], device="cuda")
        alpha = torch.tensor([0.25], device="cuda")
        rows = blocks * n_path
        fwd_bytes = rows * 128 * es * 2 + blocks * n_link * 128 * es + csr.num_edges * 4 + rows * 4
        bwd_bytes = fwd_bytes + rows * 128 * es

        def fwd():
            ops.gin_combine(csr, x, x_self=xs, eps=eps, self_mode=ops.SELF_ADD, out=out)

        def bwd():
            ops.gin_combine(csr, x, x_self=xs, eps=eps, self_mode=ops.SELF_ADD, out=out, post=ops.PostAct(z, ops.ACT_PRELU, alpha),
                            want_ddot=True)

        t_f, t_b = timed(fwd), timed(bwd)
        print("%s  fwd %7.1f us  %5.0f GB/s   bwd(post) %7.1f us  %5.0f GB/s" % (dt, t_f, fwd_bytes / t_f / 1e3, t_b, bwd_bytes / t_b / 1e3),
              flush=True)


if __name__ == "__main__":
    main()
