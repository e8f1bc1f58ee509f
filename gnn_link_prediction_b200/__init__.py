"""gnn_link_prediction_b200 — the HeteroGIN message-passing hot path of
youssefshoeb/GNN-Link-Prediction, rebuilt for B200 (sm_100a).

Layout (only what the path needs):
  csrc/          hand-written CUDA kernels + the C ABI of include/hgin.h  -> libhgin.so
  _lib.py        ctypes binding (fails loudly when the library is missing; no CPU fallback)
  ops.py         tensor-level wrappers over the C ABI
  functional.py  autograd Functions (one per heterogeneous GIN layer, one per readout layer)
  models.py      module mirror: GINConv / GINLayer / HeteroConv / HetroGIN (reference signatures)
  train.py       the reference train step + TrainStep (fused loss, flat bucket, Adam, data parallel)
  data.py        HeteroData / Batch / DataLoader without PyG (collate semantics of the reference)
  synthetic.py   seed-pinned datanet-shaped samples (reference edge order)
"""
__version__ = "0.1.0"
