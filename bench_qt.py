"""Queueing-theory baseline throughput, run as `python bench.py --workload qt`.

Workload: the Cfg-C batch (1024 synthetic 50-node topologies: 2.5 M paths, 205 k links, 7.3 M
path->link edges), capacities / offered traffic drawn so that link utilisations span 0.1 - 3.
A "step" = the whole pre-processing call the reference makes once per sample (dataset.py:86):
CSR build (both orientations, K0) + 3 fixed-point iterations + occupancy + per-path delay.
The CPU arm is the oracle port of QTBaseline.forward (models.py:42-158) on a bounded sample of
the same topologies, one sample at a time as the reference runs it.
"""
import json
import time

import torch


def main(args):
    from gnn_link_prediction_b200.baseline import QTBaseline
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    graphs = 1024
    et = ("path", "uses", "link")
    ds = SyntheticDataset(graphs, num_topologies=16, seed=1997)
    samples = [ds[i] for i in range(graphs)]
    batch = Batch.from_data_list(samples, index_dtype=torch.int32, edge_types=[et])
    g = torch.Generator().manual_seed(0)
    n_p, n_l = batch["path"].x.shape[0], batch["link"].x.shape[0]
    P = torch.rand(n_p, 3, generator=g) * 2 + 0.1
    L = torch.rand(n_l, 1, generator=g) * 60000 + 20000
    if args.impl == "reference" or not torch.cuda.is_available():
        raise SystemExit("bench.py --workload qt times the CUDA path; its CPU arm is the cpu_baseline key")
    qt = QTBaseline()
    ei, Pd, Ld = batch[et].edge_index.cuda(), P.cuda(), L.cuda()
    for _ in range(max(args.warmup, 3)):
        out = qt.forward_hetero(ei, Pd, Ld)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        out = qt.forward_hetero(ei, Pd, Ld)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    # e2e: pinned host inputs -> H2D -> call -> D2H of both results, every step
    eih, Ph, Lh = batch[et].edge_index.pin_memory(), P.pin_memory(), L.pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d, l = qt.forward_hetero(eih.cuda(non_blocking=True), Ph.cuda(non_blocking=True), Lh.cuda(non_blocking=True))
        d_h, l_h = d.cpu(), l.cpu()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import qt_oracle
        k = 16
        ptr_p, ptr_l = batch["path"].ptr, batch["link"].ptr
        t0 = time.perf_counter()
        for i in range(k):
            s = samples[i]
            qt_oracle.qt_baseline(s[et]["edge_index"], P[ptr_p[i]:ptr_p[i + 1]], L[ptr_l[i]:ptr_l[i + 1]])
        cpu_s = time.perf_counter() - t0
        cpu = {"value": k / cpu_s, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{k} of the {graphs} topologies, one sample per call as dataset.py:86 does"}
    e = batch[et].edge_index.shape[1]
    # bytes one call must move: the edge list (K0, twice), per-edge traffic 3 x (write + read), per-path / per-link scalars
    alg = e * 8 * 2 + 3 * (e * 4 * 2 + e * 4 * 2) + e * 8 + n_p * 8 + n_l * 24
    print(json.dumps({
        "metric": "QT baseline graphs/sec", "value": graphs / (ms * 1e-3), "unit": "graphs/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "qt", "desc": "queueing-theory baseline on the Cfg-C batch (1024 topologies)", "paths": n_p,
                   "links": n_l, "edges": e, "iterations": 3},
        "e2e": {"value": graphs / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": eih.numel() * 4 + Ph.numel() * 4 + Lh.numel() * 4,
                "d2h_bytes_per_step": d_h.numel() * 4 + l_h.numel() * 4},
        "gpu_launches": args.steps * (2 * 8 + 2 + 2 * 3), "cpu_baseline": cpu,
        "roofline": {"bound": "hbm", "kernel": "hgin_qt_baseline + hgin_csr_build", "achieved": alg / (ms * 1e-3) / 1e9,
                     "unit": "GB/s", "peak": 6456.2, "frac": alg / (ms * 1e-3) / 1e9 / 6456.2, "traffic": None,
                     "note": "latency/launch-bound scalar work (~0.3 GB per call): the HBM fraction is reported, not claimed"},
    }), flush=True)
