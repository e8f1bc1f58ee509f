#!/usr/bin/env python
"""Error statistics of the reduced-precision math modes against the CPU oracle on a Cfg-C-shaped batch
(hidden 128, 4 GIN layers): per tensor max |err| / max |ref|, relative Frobenius error, and the share of entries
outside rtol 1e-2 + atol 1e-2 * max|ref|.    python tools/precision_report.py [topologies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import hgin_oracle  # noqa: E402
from gnn_link_prediction_b200 import ops  # noqa: E402
from gnn_link_prediction_b200.arena import DeviceDataset, SampleArena  # noqa: E402
from gnn_link_prediction_b200.data import Batch  # noqa: E402
from gnn_link_prediction_b200.models import HetroGIN  # noqa: E402
from gnn_link_prediction_b200.synthetic import SyntheticDataset  # noqa: E402
from gnn_link_prediction_b200.train import TrainStep  # noqa: E402

KW = dict(node_embedding_size=128, message_passing_layers=4, dropout=0.0, concat_path=True, bl_features=False,
          divided_features=False, global_feats=False, mlp_layers=[128, 32], act="torch.nn.PReLU()", mlp_head_act=None,
          mlp_bn=False)


def stats(got, want):
    got, want = got.double().cpu().reshape(-1), want.double().cpu().reshape(-1)
    err = (got - want).abs()
    mx = float(want.abs().max()) + 1e-300
    bad = float((err > 1e-2 * want.abs() + 1e-2 * mx).double().mean())
    return float(err.max()) / mx, float(err.norm() / (want.norm() + 1e-300)), bad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ds = SyntheticDataset(n, num_topologies=16, seed=1997)
    samples = [ds[i] for i in range(n)]
    host = Batch.from_data_list(samples)
    torch.manual_seed(1997)
    ref = hgin_oracle.HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **KW)
    y = host["path"].y.reshape(-1, 1)
    out_ref = ref(host.x_dict, host.edge_index_dict, None)
    torch.sqrt(hgin_oracle.mape(out_ref, y)).backward()
    g_ref = {k: p.grad for k, p in ref.named_parameters()}
    dev = DeviceDataset(SampleArena.from_samples(samples, keep_coo=False))
    batch = dev.collate(list(range(n)))
    for name in ("tf32", "bf16"):
        model = HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **KW)
        model.load_state_dict(ref.state_dict())
        model.cuda().train().set_math_mode(getattr(ops, "MATH_" + name.upper()))
        step = TrainStep(model, lr=1e-3)
        with torch.no_grad():
            out = model.eval()(batch.x_dict, batch.graph, None)
        model.train()
        step(batch)
        torch.cuda.synchronize()
        print(f"== {name}: scores  max/max {stats(out, out_ref.detach())[0]:.2e}  fro {stats(out, out_ref.detach())[1]:.2e}  "
              f"outside-1e-2 {stats(out, out_ref.detach())[2]:.2e}")
        worst = (0, "")
        for k, p in model.named_parameters():
            if p.grad is None or g_ref[k] is None or p.numel() == 1:
                continue
            s = stats(p.grad, g_ref[k])
            worst = max(worst, (s[0], k))
            print(f"   {k:55s} max/max {s[0]:.2e}  fro {s[1]:.2e}  outside {s[2]:.2e}")
        print(f"   worst gradient: {worst}")


if __name__ == "__main__":
    main()
