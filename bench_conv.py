"""Conv-only roofline sweep (BASELINE.json configs[3], SURVEY §8(d) Cfg-D), run as
`python bench.py --workload cfgD`.

One synthetic heterogeneous graph, 1 M nodes (path/link/node = 900 000 / 80 000 / 20 000), 20 M
edges in 3 relations (path->link 9.96 M, link->path 9.96 M, node->link 80 000), endpoints i.i.d.
uniform (seed 0) so that gathers from the 461 MB path table are genuinely HBM-resident; F = 128
fp32.  A "step" = GIN aggregation + (1+eps) self term forward for the three relations and the
transposed-CSR gather backward, no MLP (models.py:208-215 and its autograd).  value = bytes of
SURVEY §8(d) / device time, counted per launch as that section prescribes: ALGORITHMIC (per-edge)
bytes where the gathered table exceeds L2 (the two launches that gather from the 461 MB path-sized
table), COMPULSORY bytes for the four whose table is L2-resident; the purely algorithmic figure and
the ncu DRAM traffic are reported beside it.  A block-diagonal "datanet-locality" variant (1024
topologies, the Cfg-C batch) is reported with compulsory bytes.
"""
import json
import os
import time

import torch

RELS = (("path", "uses", "link"), ("link", "includes", "path"), ("node", "has", "link"))


def uniform_graph(f, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = {"path": 900_000, "link": 80_000, "node": 20_000}
    e = {RELS[0]: 9_960_000, RELS[1]: 9_960_000, RELS[2]: 80_000}
    ei = {et: torch.stack([torch.randint(0, n[et[0]], (cnt,), generator=g, dtype=torch.int32),
                           torch.randint(0, n[et[2]], (cnt,), generator=g, dtype=torch.int32)])
          for et, cnt in e.items()}
    x = {t: torch.randn(c, f, generator=g) for t, c in n.items()}
    return n, ei, x


def datanet_graph(f, graphs=1024):
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    ds = SyntheticDataset(graphs, num_topologies=16)
    b = Batch.from_data_list([ds[i] for i in range(graphs)], index_dtype=torch.int32, edge_types=RELS)
    n = {t: b[t].x.shape[0] for t in ("path", "link", "node")}
    g = torch.Generator().manual_seed(1)
    return n, {et: b[et].edge_index for et in RELS}, {t: torch.randn(c, f, generator=g) for t, c in n.items()}


def run_variant(name, n, ei, x, steps, warmup, use_alg_bytes):
    from gnn_link_prediction_b200 import ops
    from gnn_link_prediction_b200.profiling import KernelTimer, combine_bytes
    xd = {t: v.cuda() for t, v in x.items()}
    eps = torch.full((1,), 0.1, device="cuda")
    t0 = time.perf_counter()
    fwd = {et: ops.csr_build(e.cuda(), n[et[0]], n[et[2]], by="dst") for et, e in ei.items()}
    bwd = {et: ops.csr_build(e.cuda(), n[et[0]], n[et[2]], by="src") for et, e in ei.items()}
    torch.cuda.synchronize()
    csr_s = time.perf_counter() - t0
    f = xd["path"].shape[1]
    outs = {et: torch.empty(n[et[2]], f, device="cuda") for et in RELS}
    grads = {et: torch.empty(n[et[0]], f, device="cuda") for et in RELS}
    gout = {t: torch.randn(c, f, device="cuda") for t, c in n.items()}

    def step():
        for et in RELS:   # forward: h = agg + (1+eps) x_dst
            ops.gin_combine(fwd[et], xd[et[0]], xd[et[2]], eps, ops.SELF_ADD, out=outs[et])
        for et in RELS:   # backward: dx_src = A^T dh
            ops.gin_combine(bwd[et], gout[et[2]], out=grads[et])

    from gnn_link_prediction_b200.profiling import L2_BYTES
    alg = comp = mixed = 0
    per_rel = {}
    for et in RELS:
        a1, c1 = combine_bytes(n[et[2]], n[et[0]], ei[et].shape[1], f, f, f)   # forward: gathers from the src table
        a2, c2 = combine_bytes(n[et[0]], n[et[2]], ei[et].shape[1], f, 0, f)   # backward: gathers from the dst-sized grad
        alg += a1 + a2
        comp += c1 + c2
        # per-edge (algorithmic) bytes only where the gathered table exceeds L2, else compulsory
        mixed += (a1 if n[et[0]] * f * 4 > L2_BYTES else c1) + (a2 if n[et[2]] * f * 4 > L2_BYTES else c2)
        per_rel["__".join(et)] = {"fwd_alg_bytes": a1, "bwd_alg_bytes": a2}
    for _ in range(warmup):
        step()
    timer = KernelTimer()
    torch.cuda.synchronize()
    ops.TIMER = timer
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ops.TIMER = None
    ms = a.elapsed_time(b) / steps
    # per-launch times in launch order: 3 fwd + 3 bwd per step
    per_launch = [r[1].elapsed_time(r[2]) for r in timer.records]
    names = [f"fwd {'__'.join(et)}" for et in RELS] + [f"bwd {'__'.join(et)}" for et in RELS]
    launch_ms = {nm: sum(per_launch[i::6]) / steps for i, nm in enumerate(names)}
    nbytes = mixed if use_alg_bytes else comp
    return {"variant": name, "ms_per_step": ms, "GBs": nbytes / (ms * 1e-3) / 1e9, "roofline_bytes": nbytes,
            "alg_GBs": alg / (ms * 1e-3) / 1e9, "compulsory_GBs": comp / (ms * 1e-3) / 1e9,
            "alg_bytes": alg, "compulsory_bytes": comp, "launch_ms": launch_ms, "csr_build_s": csr_s,
            "edges": sum(e.shape[1] for e in ei.values()), "nodes": sum(n.values())}


def cpu_reference(n, ei, x, frac=0.1):
    """The reference CPU path for the same op (index_select + scatter_add_ and its autograd) on a
    bounded sample: the first `frac` of each relation's edges."""
    t_total, edges = 0.0, 0
    for et in RELS:
        e = ei[et][:, : int(ei[et].shape[1] * frac)].long()
        xs = x[et[0]].clone().requires_grad_(True)
        xdst = x[et[2]]
        t0 = time.perf_counter()
        msg = xs.index_select(0, e[0])
        out = torch.zeros(n[et[2]], xs.shape[1]).scatter_add_(0, e[1].view(-1, 1).expand_as(msg), msg)
        out += (1 + 0.1) * xdst
        out.backward(torch.ones_like(out))
        t_total += time.perf_counter() - t0
        edges += e.shape[1]
    return t_total, edges


def main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    f = 128
    n, ei, x = uniform_graph(f)
    edges = sum(e.shape[1] for e in ei.values())
    if args.impl == "reference":
        t, e_used = cpu_reference(n, ei, x)
        line = {"impl": "reference", "metric": "HeteroGIN conv edges/sec (agg+self fwd, transposed bwd)",
                "value": e_used / t, "unit": "edges/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": {"workload": "cfgD", "sample": "first 10% of edges"},
                "cpu_baseline": {"value": e_used / t, "unit": "edges/s", "cores": torch.get_num_threads(),
                                 "kind": "port", "sample": "first 10% of each relation's edges, 1 pass"},
                "e2e": {"value": e_used / t, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    from bench import ClockSampler, measured_peaks
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    steps, warmup = max(args.steps, 5), max(args.warmup, 3)
    sampler = ClockSampler(0)
    sampler.start()
    uni = run_variant("uniform: algorithmic bytes where the gathered table > L2, compulsory elsewhere", n, ei, x, steps,
                      warmup, True)
    clocks = sampler.result()
    dn = run_variant("datanet block-diagonal (1024 topologies): compulsory bytes", *datanet_graph(f), steps, warmup, False)
    hbm, _, basis = measured_peaks()
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("cfgD", {}).get("gin_combine_dram_bytes_per_step")
    cpu = None
    if not args.no_cpu_baseline:
        t, e_used = cpu_reference(n, ei, x)
        cpu = {"value": e_used / t, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "first 10% of each relation's edges, fwd+bwd, 1 pass"}
    line = {"metric": "HeteroGIN conv HBM GB/s (agg+self fwd, transposed bwd, F=128)", "value": uni["GBs"],
            "unit": "GB/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": uni["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfgD", "nodes": uni["nodes"], "edges": uni["edges"], "relations": 3, "F": f,
                       "l2": "source tables larger than L2 (path table 461 MB)"},
            "edges_per_s": 2 * edges / (uni["ms_per_step"] * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "hgin_gin_combine", "achieved": uni["GBs"], "peak": hbm,
                         "unit": "GB/s", "frac": uni["GBs"] / hbm, "traffic": traffic, "peak_basis": basis,
                         "bytes": "per step (6 launches): algorithmic bytes for the two launches whose gathered "
                                  "table (461 MB) exceeds L2, compulsory bytes for the four L2-resident ones; "
                                  "traffic = ncu dram bytes per step; pure algorithmic GB/s in variants[0].alg_GBs"},
            "variants": [uni, dn], "cpu_baseline": cpu, "clocks": clocks, "gpu_launches": 6 * steps,
            "e2e": None}
    print(json.dumps(line), flush=True)
