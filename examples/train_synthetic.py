#!/usr/bin/env python
"""The reference's training flow (main.py -> train.py:163-211) on synthetic datanet-shaped samples,
with this package swapped in for `models` / the PyG loader.  Two equivalent ways to run a step:

    --mode dropin   train.py's own step body: model(...); sqrt(mape).backward(); torch.optim step
    --mode fused    gnn_link_prediction_b200.train.TrainStep (fused loss, flat bucket, hgin Adam);
                    add --graph to replay every step as one CUDA graph, or --device-dataset to keep
                    the samples in HBM and assemble every batch on the GPU (arena.DeviceDataset)

    python examples/train_synthetic.py --epochs 3 --samples 64
    torchrun --nproc-per-node 2 examples/train_synthetic.py --mode fused      # sample-sharded
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gnn_link_prediction_b200.data import CONV_EDGE_TYPES, DataLoader, DevicePrefetcher, pack_batch  # noqa: E402
from gnn_link_prediction_b200.parallel import Communicator, shard_samples  # noqa: E402
from gnn_link_prediction_b200.synthetic import SyntheticDataset  # noqa: E402
from gnn_link_prediction_b200.train import (GraphedTrainStep, TrainStep, load_model, load_optmizer, mape,  # noqa: E402
                                            test, train_one_epoch)

CONFIG = {  # /root/reference/config.json
    "SEED": 1997, "LOSS": "mape", "OPTIMIZER": "adam", "LEARNING_RATE": 0.001, "WEIGHT_DECAY": 0,
    "NODE_EMBEDDING_SIZE": 8, "MP_LAYERS": 1, "DROPOUT": 0.0, "EPOCHS": 10, "TRAIN_BATCH_SIZE": 8,
    "VAL_BATCH_SIZE": 1, "NORMALIZE_DATASET": False, "BL_FEATURES": False, "DIVIDED_FEATURES": False,
    "MODEL": "GIN", "HEADS": 16, "CONCAT_PATH": True, "GLOBAL_FEATS": False, "MLP_LAYERS": [128, 32],
    "MLP_ACT": "torch.nn.PReLU()", "MLP_BN": False, "MLP_HEAD_ACT": None,
}


class Shard:
    def __init__(self, ds, ids):
        self.ds, self.ids = ds, ids

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        return self.ds[self.ids[i]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="dropin", choices=["dropin", "fused"])
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--device-dataset", action="store_true")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--config", default=None, help="a config.json in the reference's format")
    args = ap.parse_args()
    config = dict(CONFIG)
    if args.config:
        with open(args.config) as f:
            config.update(json.load(f))
    comm = Communicator.from_env()
    torch.manual_seed(config["SEED"])                                  # main.py:48
    full = SyntheticDataset(args.samples, num_topologies=8, seed=config["SEED"])
    train_ds = Shard(full, shard_samples(int(args.samples * 0.75), comm.rank, comm.world))
    val_ds = Shard(full, list(range(int(args.samples * 0.75), args.samples)))
    model = load_model(config, {"train": full}).cuda()                 # train.py:176-177
    if args.mode == "dropin":
        if comm.world > 1:
            raise SystemExit("--mode dropin is the single-process reference step; use --mode fused under torchrun")
        loader = DataLoader(train_ds, batch_size=config["TRAIN_BATCH_SIZE"], shuffle=True)
        opt = load_optmizer(config, model)
        for epoch in range(args.epochs):                               # train.py:187-203
            model.train()
            loss, m = train_one_epoch(epoch, mape, opt, loader, model)
            model.eval()
            val = test(epoch, mape, DataLoader(val_ds, batch_size=config["VAL_BATCH_SIZE"]), model, "Validation")
            print(f"Epoch {epoch + 1} | Train Loss {loss:.4f} | MAPE-Train {m:.4f} | Validation Loss {val:.4f}")
    else:
        loader = DataLoader(train_ds, batch_size=config["TRAIN_BATCH_SIZE"], shuffle=True, pin_memory=True,
                            index_dtype=torch.int32, edge_types=CONV_EDGE_TYPES, batch_vector=False, csr=True,
                            keep_coo=False)
        step = TrainStep(model.train(), lr=config["LEARNING_RATE"], weight_decay=config["WEIGHT_DECAY"],
                         optimizer=config["OPTIMIZER"], communicator=comm)
        run = GraphedTrainStep(step) if args.graph else step
        if args.device_dataset:
            from gnn_link_prediction_b200.arena import DeviceDataset, DeviceLoader, SampleArena
            n_train = int(args.samples * 0.75)
            dev_ds = DeviceDataset(SampleArena.from_samples([full[i] for i in range(n_train)], keep_coo=False))
            loader = DeviceLoader(dev_ds, batch_size=config["TRAIN_BATCH_SIZE"], shuffle=True, rank=comm.rank,
                                  world=comm.world)
        from gnn_link_prediction_b200.train import LossReadback
        for epoch in range(args.epochs):
            total, n = 0.0, 0
            if args.device_dataset:
                batches = loader
            else:
                batches = (pack_batch(b) for b in loader) if args.graph else DevicePrefetcher(loader)
            reader = LossReadback()
            for batch in batches:
                done = reader.push(run(batch))      # the previous step's loss: no stall in the launch queue
                total += float(done[0]) if done is not None else 0.0
                n += 1
            total += float(reader.flush()[0])
            if comm.rank == 0:
                print(f"Epoch {epoch + 1} | Train Loss {total / max(n, 1):.4f} ({n} steps, {comm.world} rank(s))")


if __name__ == "__main__":
    main()
