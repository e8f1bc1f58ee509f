"""Sample-sharded data parallelism for the HeteroGIN train step.

The reference is single-process (no torch.distributed call anywhere); this is the new
multi-GPU layer.  Topology samples are independent connected components of the batched graph
(PyG collate is block-diagonal), so message passing needs NO inter-GPU exchange: each rank takes
its own samples, builds its own CSR and runs the full model.  Two collectives per step make the
result equal to the single-process step on the concatenated batch:

1. `all_reduce(sum)` of the 2 loss statistics `(sum |err/y|, N_path)` BEFORE backward: the loss
   `sqrt(100 * S / N)` (train.py:13,42) is not linear in the batch, so every rank must
   differentiate the same GLOBAL value;
2. `all_reduce(sum)` of ONE flat fp32 gradient bucket — SUM, not mean: with the global N already
   inside the seed gradient, per-rank gradients are partial sums of the global gradient.

Plumbing only (torch.distributed over NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_samples(num_samples: int, rank: int, world: int):
    """Rank r of W takes samples r, r+W, ... of the global batch.  Every rank must hold at least one
    sample (all ranks enter both all-reduces of the step): fewer samples than ranks is an error here;
    the loaders (arena.shard_id_batches) drop such a tail on all ranks instead."""
    if num_samples < world:
        raise ValueError(f"shard_samples: {num_samples} samples cannot be sharded over {world} ranks "
                         "(a rank with an empty shard would skip the step's all-reduces)")
    return list(range(rank, num_samples, world))


class Communicator:
    """Thin wrapper: a no-op when there is one process."""

    def __init__(self, group=None, enabled=None):
        self.enabled = (dist.is_available() and dist.is_initialized()) if enabled is None else enabled
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    @classmethod
    def from_env(cls, backend=None):
        """Join the job torchrun started (RANK / WORLD_SIZE / MASTER_* in the environment)."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1 and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
        return cls()

    def all_reduce_sum_(self, t):
        if self.enabled and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max_(self, t):
        if self.enabled and self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        if self.enabled and self.world > 1:
            dist.barrier(group=self.group)


def sqrt_mape_seed(pred, y, sums):
    """d sqrt(100*S/N) / d pred for GLOBAL (S, N) = sums — the formula hgin_sqrt_mape_bwd evaluates
    on the GPU, restated with torch ops for the CPU (gloo) protocol tests."""
    S, N = sums[0], sums[1]
    L = torch.sqrt(100.0 * S / N)
    u = (pred - y) / y
    return 50.0 * torch.sign(u) / (y * N * L)
