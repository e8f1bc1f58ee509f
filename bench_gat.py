"""HetroGAT train-step throughput, run as `python bench.py --workload gat`.

Workload: the Cfg-C batch (1024 synthetic 50-node topologies per step) through config.json's model with MODEL = GAT
(16 heads x 8 channels, one layer, readout 131 -> 128 -> 32 -> 1), tf32 dense layers, fp32 attention kernels.
A "step" = forward + sqrt(MAPE) + backward + Adam (TrainStep) on a batch resident in HBM; `e2e` assembles a fresh random batch on
the GPU from an HBM-resident dataset every step (arena.DeviceDataset) and reads the loss back.  The CPU arm is the oracle
port of the reference's HetroGAT on the restated PyG GATConv (parity unpinned for the PyG part) on a bounded sample.
"""
import json
import time

import torch


def main(args):
    from gnn_link_prediction_b200 import ops
    from gnn_link_prediction_b200.arena import DeviceDataset, SampleArena
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.models import HetroGAT
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    from gnn_link_prediction_b200.train import LossReadback, TrainStep
    if args.impl == "reference" or not torch.cuda.is_available():
        raise SystemExit("bench.py --workload gat times the CUDA path; its CPU arm is the cpu_baseline key")
    graphs = 1024
    kw = dict(node_embedding_size=8, message_passing_layers=1, dropout=0.0, heads=16, concat_path=True, bl_features=False,
              divided_features=False, global_feats=False, mlp_layers=[128, 32], act="torch.nn.PReLU()", mlp_head_act=None,
              mlp_bn=False)
    ds = SyntheticDataset(graphs, num_topologies=16, seed=1997)
    samples = [ds[i] for i in range(graphs)]
    dev = DeviceDataset(SampleArena.from_samples(samples, keep_coo=False))
    batch = dev.collate(list(range(graphs)))
    edges = sum(batch[et]["csr_dst_col"].shape[0] for et in batch.edge_types)
    torch.manual_seed(1997)
    model = HetroGAT(input_channels={"link": 7, "path": 7, "node": 3}, **kw).cuda().train()
    model.set_math_mode(ops.MATH_TF32)
    step = TrainStep(model)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(batch)
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(batch)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    # end to end: ids from pinned host memory -> on-GPU collate of a fresh random batch -> step -> loss read-back
    gen = torch.Generator().manual_seed(7)
    ids = [torch.randint(0, graphs, (graphs,), generator=gen).numpy() for _ in range(args.steps + 2)]
    reader = LossReadback()
    for i in ids[:2]:
        reader.push(step(dev.collate(i)))
    reader.flush()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in ids[2:]:
        reader.push(step(dev.collate(i)))
    last = reader.flush()
    b.record()
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(b) / args.steps

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import hgin_oracle
        n_cpu = 64
        host = Batch.from_data_list(samples[:n_cpu])
        torch.manual_seed(1997)
        ref = hgin_oracle.HetroGAT(input_channels={"link": 7, "path": 7, "node": 3}, **kw)
        opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
        hgin_oracle.train_step(ref, opt, host)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            hgin_oracle.train_step(ref, opt, host)
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_cpu / dt, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n_cpu} topologies/step x {reps} steps, fwd+bwd+Adam, oracle port of HetroGAT on the restated GATConv"}
    line = {"metric": "HetroGAT train graphs/sec", "value": graphs / (ms * 1e-3), "unit": "graphs/s", "n_gpus": 1, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (dense layers; f32 attention, accumulation)", "data": "synthetic",
            "config": {"workload": "gat", "desc": "config.json with MODEL = GAT (16 heads x 8, 1 layer), 1024 topologies/step",
                       "graphs_per_gpu_per_step": graphs, "edges_per_gpu_per_step": edges,
                       "l2": "inputs+activations larger than L2"},
            "e2e": {"value": graphs / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": dev.h2d_bytes(graphs), "d2h_bytes_per_step": 8,
                    "collate": "inside the timed region, on the GPU (arena.DeviceDataset)"},
            "gpu_launches": None, "roofline": None, "cpu_baseline": cpu, "loss_last": float(last[0])}
    print(json.dumps(line), flush=True)
