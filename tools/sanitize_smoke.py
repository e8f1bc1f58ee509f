"""Small invocation of every kernel family (for compute-sanitizer memcheck runs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_link_prediction_b200 import ops  # noqa: E402
from gnn_link_prediction_b200.data import Batch  # noqa: E402
from gnn_link_prediction_b200.models import HetroGIN, MATH_TF32  # noqa: E402
from gnn_link_prediction_b200.synthetic import SyntheticDataset  # noqa: E402
from gnn_link_prediction_b200.train import TrainStep  # noqa: E402

torch.manual_seed(0)
for emb, layers, math in [(8, 1, None), (128, 3, MATH_TF32), (32, 2, None)]:
    ds = SyntheticDataset(3, num_nodes=14, num_links=24, num_topologies=2)
    b = Batch.from_data_list([ds[i] for i in range(3)], index_dtype=torch.int32).cuda()
    m = HetroGIN({"link": 7, "path": 7, "node": 3}, emb, layers, 0.0, True, False, False, False, [128, 32],
                 "torch.nn.PReLU()", None, False).cuda().train()
    if math is not None:
        m.set_math_mode(math)
    step = TrainStep(m)
    for _ in range(2):
        loss = step(b)
    torch.cuda.synchronize()
    print("emb", emb, "layers", layers, "loss", float(loss[0]))
# ragged / tiny shapes through the raw ops
ei = torch.tensor([[0, 2, 2, 1, 0], [1, 1, 0, 1, 1]], device="cuda")
csr = ops.csr_build(ei, 3, 2).validate()
x = torch.randn(3, 5, device="cuda")
print(ops.gin_combine(csr, x).sum().item())
a, bb = torch.randn(300, 128, device="cuda"), torch.randn(300, 64, device="cuda")
print(float(ops.debug_gemm_tn(a, bb).abs().sum()))
print("sanitize smoke done")
