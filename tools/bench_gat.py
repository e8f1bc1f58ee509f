#!/usr/bin/env python
"""HetroGAT train step (config.json with MODEL = GAT: 16 heads x 8 channels, one layer) on a Cfg-C-sized batch:
ms per step and the per-op breakdown.    python tools/bench_gat.py [topologies] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gnn_link_prediction_b200 import ops  # noqa: E402
from gnn_link_prediction_b200.arena import DeviceDataset, SampleArena  # noqa: E402
from gnn_link_prediction_b200.models import HetroGAT  # noqa: E402
from gnn_link_prediction_b200.profiling import KernelTimer  # noqa: E402
from gnn_link_prediction_b200.synthetic import SyntheticDataset  # noqa: E402
from gnn_link_prediction_b200.train import TrainStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ds = SyntheticDataset(n, num_topologies=16, seed=1997)
dev = DeviceDataset(SampleArena.from_samples([ds[i] for i in range(n)], keep_coo=False))
batch = dev.collate(list(range(n)))
torch.manual_seed(1997)
model = HetroGAT(input_channels={"link": 7, "path": 7, "node": 3}, node_embedding_size=8, message_passing_layers=1, dropout=0.0,
                 heads=16, concat_path=True, bl_features=False, divided_features=False, global_feats=False,
                 mlp_layers=[128, 32], act="torch.nn.PReLU()", mlp_head_act=None, mlp_bn=False).cuda().train()
model.set_math_mode(ops.MATH_TF32)
step = TrainStep(model)
for _ in range(3):
    step(batch)
torch.cuda.synchronize()
timer = KernelTimer()
ops.TIMER = timer
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    loss = step(batch)
b.record()
torch.cuda.synchronize()
ops.TIMER = None
ms = a.elapsed_time(b) / steps
print(f"HetroGAT, {n} topologies/step: {ms:.2f} ms/step = {n / ms * 1e3:.0f} graphs/s, loss {float(loss[0]):.3f}")
for k, v in timer.summary().items():
    print(f"   {k:14s} {v['ms'] / steps:8.3f} ms/step  {v['launches'] / steps:5.1f} calls")
