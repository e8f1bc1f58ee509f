import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


# the non-default branches of HetroGIN (models.py:301-330, 347-371): global_feats, mlp_bn, other activations
FLAG_CASES = ["L2_emb8_globalfeats", "L1_emb8_globalfeats_noconcat", "L2_emb8_bn", "L1_emb8_bn_leaky_headrelu",
              "L2_emb8_elu_softplus", "L1_emb8_gelu", "L1_emb8_tanh_headsigmoid", "L1_emb8_silu"]
# HetroGAT (models.py:380-506) on the restated PyG GATConv
GAT_CASES = ["gat_default", "gat_h4_emb16_noconcat", "gat_h1_L2_emb8", "gat_h2_emb4_bl_globalfeats"]
MODEL_CASES = (["default", "L3_emb16", "L2_emb32_noconcat", "L2_emb8_blfeat", "L2_emb12_divided_bl"] + FLAG_CASES
               + GAT_CASES)


def build_model(module, cfg, input_channels):
    """`module.HetroGIN` / `module.HetroGAT` (the package's models or the oracle's) for a config.json dict, as
    train.py:116-137 builds them."""
    kw = config_to_kwargs(cfg)
    if cfg["MODEL"] == "GAT":
        return module.HetroGAT(input_channels=input_channels, heads=cfg["HEADS"], **kw)
    return module.HetroGIN(input_channels=input_channels, **kw)


def config_to_kwargs(cfg):
    """config.json keys -> HetroGIN constructor kwargs, as train.py:128-132 maps them."""
    return dict(node_embedding_size=cfg["NODE_EMBEDDING_SIZE"], message_passing_layers=cfg["MP_LAYERS"],
                dropout=cfg["DROPOUT"], concat_path=cfg["CONCAT_PATH"], bl_features=cfg["BL_FEATURES"],
                divided_features=cfg["DIVIDED_FEATURES"], global_feats=cfg["GLOBAL_FEATS"],
                mlp_layers=cfg["MLP_LAYERS"], act=cfg["MLP_ACT"], mlp_bn=cfg["MLP_BN"],
                mlp_head_act=cfg["MLP_HEAD_ACT"])
