"""Probe of the staged long-row kernel (csrc/gin_stage_blocks.cuh) on a Cfg-C shaped relation: per-launch time of the
default gather kernel, of the staged kernel, and of the staged kernel with parts switched off (HGIN_SG_DEBUG)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_link_prediction_b200 import ops                      # noqa: E402
from gnn_link_prediction_b200.functional import GraphCSR      # noqa: E402


def relation(blocks, n_in, n_out, hops, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.arange(blocks * n_in).repeat_interleave(hops)
    blk = src // n_in
    dst = torch.randint(0, n_out, (src.numel(),), generator=g) + blk * n_out
    ptr_in = torch.arange(blocks + 1, dtype=torch.int64) * n_in
    ptr_out = torch.arange(blocks + 1, dtype=torch.int64) * n_out
    return torch.stack((src, dst)), ptr_in, ptr_out


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    blocks, n_in, n_out, hops = 1024, 2450, 200, 3
    ei, ptr_in, ptr_out = relation(blocks, n_in, n_out, hops, 0)
    et = ("path", "to", "link")
    for dt in (torch.float32, torch.bfloat16):
        ops.STAGE_LONG_ROWS = True
        graph = GraphCSR({et: ei.cuda()}, {"path": blocks * n_in, "link": blocks * n_out},
                         blocks={"path": ptr_in.cuda(), "link": ptr_out.cuda()})
        plan = graph.stream_plan(et, "fwd")
        x = torch.randn(blocks * n_in, 128, device="cuda").to(dt)
        xs = torch.randn(blocks * n_out, 128, device="cuda").to(dt)
        eps = torch.tensor([0.1], device="cuda")
        out = torch.empty(blocks * n_out, 128, device="cuda", dtype=dt)
        kw = dict(x_self=xs, eps=eps, self_mode=ops.SELF_ADD, out=out)
        print(dt, "gate", plan.gate.tolist(), flush=True)
        print("  gather kernel          %8.1f us" % timed(lambda: ops.gin_combine(graph.fwd(et), x, **kw)), flush=True)
        for dbg, name in ((0, "staged"), (1, "staged, no gathers"), (2, "staged, no copies"), (3, "staged, neither"),
                          (4, "staged, 16 KB copies"), (5, "staged, 16 KB copies, no gathers")):
            os.environ["HGIN_SG_DEBUG"] = str(dbg)
            print("  %-32s %8.1f us" % (name, timed(lambda: ops.gin_combine(graph.fwd(et), x, block_plan=plan, **kw))), flush=True)
        os.environ.pop("HGIN_SG_DEBUG")


if __name__ == "__main__":
    main()
