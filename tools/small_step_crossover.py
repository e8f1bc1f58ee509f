"""hgin_small_step against the layer-by-layer TrainStep (fp32 SIMT and tf32 tensor-core modes) for config.json's model at
batch sizes from 8 to 1024 topologies (eager launches).    python tools/small_step_crossover.py"""
import sys, torch
sys.path.insert(0, ".")
from gnn_link_prediction_b200.arena import DeviceDataset, SampleArena
from gnn_link_prediction_b200.models import HetroGIN
from gnn_link_prediction_b200.synthetic import SyntheticDataset
from gnn_link_prediction_b200.train import TrainStep
from gnn_link_prediction_b200 import ops
kw = dict(node_embedding_size=8, message_passing_layers=1, dropout=0.0, concat_path=True, bl_features=False, divided_features=False,
          global_feats=False, mlp_layers=[128, 32], act="torch.nn.PReLU()", mlp_head_act=None, mlp_bn=False)
ds = SyntheticDataset(1024, num_topologies=16, seed=1)
dev = DeviceDataset(SampleArena.from_samples([ds[i] for i in range(1024)], keep_coo=False))
for n in (8, 32, 128, 512, 1024):
    batch = dev.collate(list(range(n)))
    res = []
    for fused, mode in ((True, ops.MATH_FP32), (False, ops.MATH_FP32), (False, ops.MATH_TF32)):
        torch.manual_seed(0)
        m = HetroGIN({"link": 7, "path": 7, "node": 3}, **kw).cuda().train()
        m.set_math_mode(mode)
        step = TrainStep(m, fused_small=fused)
        for _ in range(3): step(batch)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): step(batch)
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 10)
    print(f"{n:5d} topologies ({n*2450} path rows): small_step {res[0]:.3f} ms | layer-by-layer fp32 {res[1]:.3f} ms | tf32 {res[2]:.3f} ms")
