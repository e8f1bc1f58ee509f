"""TEST INFRASTRUCTURE — golden vectors for the queueing-theory baseline.

Runs only in the build container (needs /root/reference):  python oracle/make_golden_qt.py

Fabricated (topology, routing, traffic, capacity) inputs go through the reference's OWN pipeline —
`simulation_to_networkX` + `from_networkx` (generateFiles.py:21-190), the `data.type` line of
`process_file` (generateFiles.py:227-229) and the P / L assembly of `GNN21Dataset.preprocess`
(dataset.py:66-83) — and then through the UNMODIFIED `QTBaseline.forward` (models.py:42-158) on
`oracle/pyg_shim`.  Inputs and outputs are recorded in tests/golden/qt_*.pt.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "pyg_shim"), REF, ROOT]
os.environ.setdefault("WANDB_MODE", "disabled")

import numpy as np  # noqa: E402
import networkx as nx  # noqa: E402
import torch  # noqa: E402

import generateFiles as ref_gen  # noqa: E402  (reference, unmodified)
import models as ref_models  # noqa: E402  (reference, unmodified)

from gnn_link_prediction_b200.synthetic import Topology  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# separate_edge_timesteps (models.py:19-26) relies on torch<=1.9 semantics: `torch.clip(long_tensor, 0., 1.)`
# kept the integer dtype, so the index_put on the next line matched dtypes.  torch 2.x promotes to float and
# the UNMODIFIED reference raises "Index put requires the source and destination dtypes match".  The
# reference file stays untouched; the old clamp semantics are restored around the call instead.
_ORIG_CLIP = torch.clip


def _clip_keeping_integer_dtype(x, lo=None, hi=None):
    if not x.is_floating_point():
        return _ORIG_CLIP(x, None if lo is None else int(lo), None if hi is None else int(hi))
    return _ORIG_CLIP(x, lo, hi)


def reference_data(spec, seed, load):
    """Homogeneous `Data` exactly as generateFiles.process_file + dataset.preprocess build it."""
    topo = Topology(*spec)
    rng = np.random.RandomState(seed)
    n = len(topo.adj)
    G = nx.DiGraph()
    for i in range(n):
        G.add_node(i, queueSizes=32, levelsQoS=1, schedulingPolicy=0)
    for a in range(n):
        for b in topo.adj[a]:
            G.add_edge(a, b, bandwidth=float(rng.choice([10000.0, 25000.0, 40000.0])))
    R = np.empty((n, n), dtype=object)
    T = np.empty((n, n), dtype=object)
    D = np.empty((n, n), dtype=object)
    for s in range(n):
        for d in range(n):
            R[s, d] = topo.routes[s][d]
            T[s, d] = {"Flows": [{"AvgBw": float(rng.uniform(100.0, 2000.0)), "PktsGen": float(rng.uniform(0.2, 2.0) * load),
                                  "ToS": 0, "SizeDistParams": {"AvgPktSize": 1.0},
                                  "TimeDistParams": {"AvgPktsLambda": float(rng.uniform(0.2, 2.0))}}]}
            D[s, d] = {"Flows": [{"AvgDelay": float(rng.uniform(0.1, 9.0))}]}
    graph = ref_gen.simulation_to_networkX(G, R, T, D, None)
    data = ref_gen.from_networkx(graph)
    data.edge_index = data.edge_index.int()                                              # generateFiles.py:227
    data.type = torch.as_tensor(np.array([ref_gen.name_to_id(name) for name in graph.nodes]))   # generateFiles.py:229
    # dataset.py:66-83
    data.p_AvgBw = data.p_AvgBw / 1000.0
    data.P = torch.cat([getattr(data, a).view(-1, 1) for a in ["p_time_AvgPktsLambda", "p_PktsGen", "p_AvgBw"]], axis=1)
    data.L = data.l_capacity.clone().view(-1, 1)
    return data


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, spec, seed, load in [("n12", (12, 20, 3), 0, 6.0), ("n8_light", (8, 7, 1), 1, 1.0), ("n20_heavy", (20, 60, 9), 2, 5.0),
                                   ("n8_overflow", (8, 7, 1), 3, 80.0)]:      # rho^33 overflows fp32: the NaN pattern is pinned too
        data = reference_data(spec, seed, load)
        inputs = {"edge_index": data.edge_index.clone(), "edge_type": data.edge_type.clone(), "type": data.type.clone(),
                  "P": data.P.float().clone(), "L": data.L.float().clone(), "p_l": data["p-l"].clone()}
        data.P, data.L = inputs["P"].clone(), inputs["L"].clone()
        torch.clip = _clip_keeping_integer_dtype
        try:
            out_paths, out_links = ref_models.QTBaseline()(data)                         # dataset.py:86
        finally:
            torch.clip = _ORIG_CLIP
        torch.save({**inputs, "spec": list(spec), "out_paths": out_paths.clone(), "out_links": out_links.clone()},
                   os.path.join(GOLDEN, f"qt_{name}.pt"))
        rho = out_links[:, 1]
        print(f"qt_{name}: paths {tuple(out_paths.shape)} links {tuple(out_links.shape)} rho [{float(rho.min()):.3f}, "
              f"{float(rho.max()):.3f}] delay mean {float(out_paths.mean()):.4f} finite {bool(torch.isfinite(out_links).all())}")


if __name__ == "__main__":
    main()
