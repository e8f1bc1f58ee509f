#!/usr/bin/env python
"""Per-tensor errors of the GPU model against one golden fixture (tests/golden/model_<case>.pt): forward output, every
gradient, and the 5-step Adam trajectory.    python tools/fixture_report.py <case> [<case> ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from conftest import config_to_kwargs, load_golden  # noqa: E402
from gnn_link_prediction_b200.models import HetroGIN  # noqa: E402
from gnn_link_prediction_b200.train import mape  # noqa: E402


def err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double()
    return float((got - want).abs().max()), float(want.abs().max())


for case in sys.argv[1:]:
    fx = load_golden(f"model_{case}.pt")
    in_ch = {k: v.shape[1] for k, v in fx["x_dict"].items()}
    m = HetroGIN(input_channels=in_ch, **config_to_kwargs(fx["config"]))
    m.load_state_dict(fx["state_dict"])
    m.cuda().train()
    cuda = lambda d: {k: v.cuda() for k, v in d.items()}
    out = m(cuda(fx["x_dict"]), cuda(fx["edge_index_dict"]), fx["path_batch"].cuda())
    print(f"== {case}: out err/max {err(out, fx['out'])}")
    torch.sqrt(mape(out, fx["y"].cuda().reshape(-1, 1))).backward()
    for k, p in m.named_parameters():
        g = fx["grads"][k]
        if g is not None:
            e, mx = err(p.grad, g)
            print(f"   grad {k:55s} err {e:.3e}  max {mx:.3e}  rel {e / (mx + 1e-300):.2e}")
    opt = torch.optim.Adam(m.parameters(), lr=fx["config"]["LEARNING_RATE"])
    m.load_state_dict(fx["state_dict"])
    losses = []
    for _ in fx["losses"]:
        opt.zero_grad()
        out = m(cuda(fx["x_dict"]), cuda(fx["edge_index_dict"]), fx["path_batch"].cuda())
        lv = mape(out, fx["y"].cuda().reshape(-1, 1))
        torch.sqrt(lv).backward()
        opt.step()
        losses.append(float(lv))
    print("   losses", losses, "\n   ref   ", fx["losses"])
    for k, v in m.state_dict().items():
        e, mx = err(v.float(), fx["final_state_dict"][k].float())
        if e > 1e-4 * mx + 1e-7:
            print(f"   final {k:55s} err {e:.3e}  max {mx:.3e}")
