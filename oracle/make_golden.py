"""TEST INFRASTRUCTURE — generates tests/golden/*.pt from the UNMODIFIED reference.

Runs only in the build container, where /root/reference exists:

    python oracle/make_golden.py

It puts `oracle/pyg_shim` (restated PyG 2.0.2 subset) and `/root/reference` on sys.path, imports
the reference's own `models.py`, `train.py` (for `mape`, `load_model`, `load_optmizer`) and
`generateFiles.py`, and records, for seed-pinned synthetic inputs:

* edges_*.pt  — the six COO relations produced by the reference's `simulation_to_networkX` +
                `from_networkx` (generateFiles.py:21-190) for a fabricated topology/routing;
* model_*.pt  — inputs, `state_dict`, forward output, loss, every parameter gradient (or None),
                and a 5-step Adam trajectory of the reference step body (train.py:31-44).

The fixtures are what tests compare against on machines that do not have /root/reference
(the GPU box).  Nothing here is imported by the product.
"""
import json
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
import train as ref_train  # noqa: E402  (reference, unmodified)
from torch_geometric.data import Batch as ShimBatch, HeteroData as ShimHeteroData  # noqa: E402

from gnn_link_prediction_b200.synthetic import Topology, make_sample  # noqa: E402
from gnn_link_prediction_b200.data import EDGE_TYPES  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REL_NAMES = {"p-l": EDGE_TYPES[0], "l-p": EDGE_TYPES[1], "l-n": EDGE_TYPES[2],
             "n-l": EDGE_TYPES[3], "p-n": EDGE_TYPES[4], "n-p": EDGE_TYPES[5]}


def reference_relations(topo):
    """Feed a fabricated (topology, R, T, D) through the reference's own graph builder."""
    n = len(topo.adj)
    G = nx.DiGraph()
    for i in range(n):
        G.add_node(i, queueSizes=32, levelsQoS=1, schedulingPolicy=0)
    for a in range(n):
        for b in topo.adj[a]:
            G.add_edge(a, b, bandwidth=10000.0)
    R = np.empty((n, n), dtype=object)
    T = np.empty((n, n), dtype=object)
    D = np.empty((n, n), dtype=object)
    for s in range(n):
        for d in range(n):
            R[s, d] = topo.routes[s][d]
            T[s, d] = {"Flows": [{"AvgBw": 1.0 + s + d, "PktsGen": 2.0, "ToS": 0,
                                  "SizeDistParams": {"AvgPktSize": 1.0},
                                  "TimeDistParams": {"AvgPktsLambda": 1.0}}]}
            D[s, d] = {"Flows": [{"AvgDelay": 0.5}]}
    data = ref_gen.from_networkx(ref_gen.simulation_to_networkX(G, R, T, D, None))
    return {et: data[name] for name, et in REL_NAMES.items()}


def to_shim(sample):
    d = ShimHeteroData()
    for nt in sample.node_types:
        for k, v in sample[nt].items():
            d[nt][k] = v.clone()
    for et in sample.edge_types:
        d[et].edge_index = sample[et].edge_index.clone()
    return d


def model_case(name, topo_specs, config, steps=5):
    torch.manual_seed(config["SEED"])
    samples = [make_sample(Topology(*spec), seed=config["SEED"] + i) for i, spec in enumerate(topo_specs)]
    batch = ShimBatch.from_data_list([to_shim(s) for s in samples])
    datasets = {"train": [to_shim(samples[0])]}
    model = ref_train.load_model(config, datasets)          # train.py:116-137
    opt = ref_train.load_optmizer(config, model)            # train.py:140-148
    model.train()
    if config["MODEL"] == "GAT":
        # GATConv((-1, -1), ...) defers its Linear weights to the first forward call (models.py:417-420): one dry run
        # materialises them (glorot draws in call order) so that the fixture can record a complete state_dict
        with torch.no_grad():
            model(batch.x_dict, batch.edge_index_dict, batch["path"].batch)
    fixture = {
        "config": dict(config),
        "topologies": [list(s) for s in topo_specs],
        "x_dict": {k: v.clone() for k, v in batch.x_dict.items()},
        "edge_index_dict": {k: v.clone() for k, v in batch.edge_index_dict.items()},
        "y": batch["path"].y.clone(),
        "path_batch": batch["path"].batch.clone(),
        "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
    }
    losses = []
    for step in range(steps):
        # -- the reference step body, train.py:31-44 (sample.cuda() dropped: CPU path) --------
        opt.zero_grad()
        out = model(batch.x_dict, batch.edge_index_dict, batch["path"].batch)
        label = batch["path"].y.reshape(-1, 1)
        loss_value = ref_train.mape(out, label)
        loss = torch.sqrt(loss_value)
        loss.backward()
        if step == 0:
            fixture["out"] = out.detach().clone()
            fixture["loss_value"] = loss_value.detach().clone()
            fixture["grads"] = {k: (None if p.grad is None else p.grad.detach().clone())
                                for k, p in model.named_parameters()}
        opt.step()
        losses.append(float(loss_value))
        # forward() rebinds x_dict['path'/'link'] on the dict it is given (models.py:334-338), but
        # batch.x_dict builds a fresh dict each access, so the stored features stay 7 columns wide.
    fixture["losses"] = losses
    fixture["final_state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
    torch.save(fixture, os.path.join(GOLDEN, f"model_{name}.pt"))
    n_none = sum(v is None for v in fixture["grads"].values())
    print(f"model_{name}: out {tuple(fixture['out'].shape)} loss {losses[0]:.6f} -> {losses[-1]:.6f}; "
          f"{len(fixture['grads'])} params, {n_none} with grad None")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for spec in [(12, 20, 3), (8, 7, 1), (20, 60, 9)]:
        topo = Topology(*spec)
        rel = reference_relations(topo)
        torch.save({"spec": list(spec), "relations": rel,
                    "n_path": topo.n_path, "n_link": topo.n_link, "n_node": topo.n_node},
                   os.path.join(GOLDEN, f"edges_n{spec[0]}.pt"))
        print(f"edges_n{spec[0]}:", {"__".join(k): tuple(v.shape) for k, v in rel.items()})

    with open(os.path.join(REF, "config.json")) as f:
        base = json.load(f)
    model_case("default", [(12, 20, 3), (10, 14, 4)], base)
    model_case("L3_emb16", [(12, 20, 3), (8, 7, 1), (10, 14, 4)],
               {**base, "NODE_EMBEDDING_SIZE": 16, "MP_LAYERS": 3, "MLP_LAYERS": [32, 16]})
    model_case("L2_emb32_noconcat", [(12, 20, 5)],
               {**base, "NODE_EMBEDDING_SIZE": 32, "MP_LAYERS": 2, "CONCAT_PATH": False, "MLP_LAYERS": [24]})
    model_case("L2_emb8_blfeat", [(10, 14, 4), (8, 7, 1)],
               {**base, "MP_LAYERS": 2, "BL_FEATURES": True})
    # DIVIDED_FEATURES=true with BL_FEATURES=false raises inside the reference itself (channel
    # arithmetic gives link 4 columns, models.py:265-269, the slicing gives 3, models.py:341), so
    # the divided variant that works is pinned: both flags on, all 7+7 columns used.
    model_case("L2_emb12_divided_bl", [(10, 14, 4)],
               {**base, "NODE_EMBEDDING_SIZE": 12, "MP_LAYERS": 2, "DIVIDED_FEATURES": True,
                "BL_FEATURES": True})
    # ---- the non-default branches of HetroGIN (models.py:301-330, 347-371) -------------------------------------------
    # GLOBAL_FEATS only works with 4 path columns (global_feats_size is the constant 8 = 2 x 4, models.py:271-274), i.e.
    # BL_FEATURES on and DIVIDED_FEATURES off; every other combination raises inside the reference's first readout Linear.
    model_case("L2_emb8_globalfeats", [(10, 14, 4), (8, 7, 1), (12, 20, 3)],
               {**base, "MP_LAYERS": 2, "BL_FEATURES": True, "GLOBAL_FEATS": True})
    model_case("L1_emb8_globalfeats_noconcat", [(10, 14, 4), (8, 7, 1)],
               {**base, "BL_FEATURES": True, "GLOBAL_FEATS": True, "CONCAT_PATH": False, "MLP_LAYERS": [16]})
    model_case("L2_emb8_bn", [(10, 14, 4), (8, 7, 1), (12, 20, 3)],
               {**base, "MP_LAYERS": 2, "MLP_BN": True, "MLP_LAYERS": [32, 16]})
    model_case("L1_emb8_bn_leaky_headrelu", [(10, 14, 4), (8, 7, 1)],
               {**base, "MLP_BN": True, "MLP_ACT": "torch.nn.LeakyReLU(0.1)", "MLP_HEAD_ACT": "torch.nn.ReLU()",
                "MLP_LAYERS": [16, 8]})
    model_case("L2_emb8_elu_softplus", [(10, 14, 4), (8, 7, 1)],
               {**base, "MP_LAYERS": 2, "MLP_ACT": "torch.nn.ELU()", "MLP_HEAD_ACT": "torch.nn.Softplus()"})
    model_case("L1_emb8_gelu", [(12, 20, 3)], {**base, "MLP_ACT": "torch.nn.GELU()"})
    model_case("L1_emb8_tanh_headsigmoid", [(12, 20, 3)],
               {**base, "MLP_ACT": "torch.nn.Tanh()", "MLP_HEAD_ACT": "torch.nn.Sigmoid()"})
    model_case("L1_emb8_silu", [(12, 20, 3)], {**base, "MLP_ACT": "torch.nn.SiLU()"})
    # ---- HetroGAT (models.py:380-506) on the restated GATConv (PARITY UNPINNED, see oracle/pyg_shim) ----------------------
    gat = {**base, "MODEL": "GAT"}
    model_case("gat_default", [(12, 20, 3), (10, 14, 4)], gat)
    model_case("gat_h4_emb16_noconcat", [(10, 14, 4), (8, 7, 1)],
               {**gat, "HEADS": 4, "NODE_EMBEDDING_SIZE": 16, "CONCAT_PATH": False, "MLP_LAYERS": [24]})
    # layers >= 1 are built for `emb` input columns while layer 0 emits emb * heads (models.py:413-428): only HEADS = 1
    # runs with MP_LAYERS > 1
    model_case("gat_h1_L2_emb8", [(12, 20, 3)], {**gat, "HEADS": 1, "MP_LAYERS": 2})
    model_case("gat_h2_emb4_bl_globalfeats", [(10, 14, 4), (8, 7, 1)],
               {**gat, "HEADS": 2, "NODE_EMBEDDING_SIZE": 4, "BL_FEATURES": True, "GLOBAL_FEATS": True})


if __name__ == "__main__":
    main()
