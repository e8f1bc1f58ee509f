"""Data-parallel protocol on CPU: 2 gloo processes, each running the ORACLE on
its shard of samples, joined by gnn_link_prediction_b200.parallel's two all-reduces, must
reproduce the single-process gradients on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from gnn_link_prediction_b200.parallel import shard_samples


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


KW = dict(node_embedding_size=16, message_passing_layers=2, dropout=0.0, concat_path=True, bl_features=False,
          divided_features=False, global_feats=False, mlp_layers=[32, 16], act="torch.nn.PReLU()", mlp_head_act=None,
          mlp_bn=False)


def _worker(rank, world, port, num_samples, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    import torch.distributed as dist
    from oracle import hgin_oracle
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.parallel import Communicator, sqrt_mape_seed
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    comm = Communicator.from_env("gloo")
    assert comm.world == world and comm.rank == rank
    ds = SyntheticDataset(num_samples, num_nodes=10, num_links=14, num_topologies=3)
    mine = shard_samples(num_samples, rank, world)
    batch = Batch.from_data_list([ds[i] for i in mine])
    torch.manual_seed(11)                                   # identical replicas
    model = hgin_oracle.HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **KW)
    out = model(batch.x_dict, batch.edge_index_dict, None)
    y = batch["path"].y.reshape(-1, 1)
    sums = torch.stack([torch.abs((out.detach() - y) / y).sum(), torch.tensor(float(y.numel()))])
    comm.all_reduce_sum_(sums)                              # collective 1: global loss statistics
    out.backward(sqrt_mape_seed(out.detach(), y, sums))
    live = [p for p in model.parameters() if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for p in live])
    comm.all_reduce_sum_(flat)                              # collective 2: SUM of the flat bucket
    comm.barrier()
    if rank == 0:
        torch.save({"flat": flat, "mape": 100.0 * sums[0] / sums[1]}, os.path.join(out_dir, "dp.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("num_samples", [4, 5])
def test_two_rank_gloo_equals_single_process(tmp_path, num_samples):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), num_samples, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "dp.pt")
    from oracle import hgin_oracle
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    ds = SyntheticDataset(num_samples, num_nodes=10, num_links=14, num_topologies=3)
    batch = Batch.from_data_list([ds[i] for i in range(num_samples)])
    torch.manual_seed(11)
    model = hgin_oracle.HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **KW)
    out = model(batch.x_dict, batch.edge_index_dict, None)
    y = batch["path"].y.reshape(-1, 1)
    loss_value = hgin_oracle.mape(out, y)
    torch.sqrt(loss_value).backward()                       # train.py:40-43 on the whole batch
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    torch.testing.assert_close(got["mape"], loss_value.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["flat"], flat, rtol=2e-4, atol=1e-7)


def test_shard_samples_partitions_the_batch():
    for n, w in [(8, 2), (5, 2), (1024, 8), (4, 4)]:
        parts = [shard_samples(n, r, w) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n)) and all(parts)
    with pytest.raises(ValueError):          # a rank with no samples would skip the step's all-reduces
        shard_samples(3, 0, 4)


def _loader_worker(rank, world, port, n, bs, out_dir):
    """Every rank iterates its shard of the loader and enters one all-reduce per step: with a ragged tail
    (n % (bs * world) == 1) a loader that yields on some ranks only would hang here."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import datetime
    import torch.distributed as dist
    from gnn_link_prediction_b200.arena import DeviceLoader
    dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=30))

    class Stub:
        def __len__(self):
            return n

        def collate(self, ids):
            return [int(i) for i in ids]

    loader = DeviceLoader(Stub(), batch_size=bs, shuffle=True, generator=torch.Generator().manual_seed(2), rank=rank,
                          world=world)
    steps, total = 0, torch.zeros(1)
    for ids in loader:
        t = torch.tensor([float(len(ids))])
        dist.all_reduce(t)
        total += t
        steps += 1
    assert steps == len(loader)
    if rank == 0:
        torch.save({"steps": steps, "total": float(total)}, os.path.join(out_dir, "loader.pt"))
    dist.destroy_process_group()


def test_two_rank_loader_with_ragged_tail_keeps_ranks_in_step(tmp_path):
    n, bs, world = 9, 2, 2                   # 9 % (2 * 2) == 1: the last chunk has one sample for two ranks
    mp.spawn(_loader_worker, args=(world, _free_port(), n, bs, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "loader.pt")
    assert got == {"steps": 2, "total": 8.0}


VARIANTS = {
    "gin": dict(),
    # mlp_bn: BatchNorm statistics must be those of the GLOBAL batch (column sums all-reduced) for N ranks to equal one;
    # global_feats: per-graph pools are local to a sample, so sharding by sample leaves them unchanged
    "gin_bn_globalfeats": dict(mlp_bn=True, global_feats=True, bl_features=True),
    # config.json's own model: the three-kernel step (hgin_small_step) split around the all-reduce of the loss statistics
    "gin_default_small_step": dict(node_embedding_size=8, message_passing_layers=1, mlp_layers=[128, 32]),
    # (HetroGAT is NOT in this list: PyG's bipartite add_self_loops rule — GATConv drops edges whose source id equals
    # their destination id and adds (i, i) for i < min(N_src, N_dst), ids taken across node types — makes its output a
    # function of the batch composition, so a sharded step cannot equal the single-process one by construction.)
}


def _build(variant):
    from gnn_link_prediction_b200.models import HetroGAT, HetroGIN
    opts = dict(VARIANTS[variant])
    graphed = opts.pop("graphed", False)
    heads = opts.pop("gat_heads", None)
    kw = {**KW, **opts}
    torch.manual_seed(11)
    if heads is not None:
        model = HetroGAT(input_channels={"link": 7, "path": 7, "node": 3}, heads=heads, **kw)
    else:
        model = HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **kw)
    return model.cuda().train(), graphed


def _gpu_worker(rank, world, port, num_samples, out_dir, variant="gin"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.models import HetroGIN
    from gnn_link_prediction_b200.parallel import Communicator
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    from gnn_link_prediction_b200.train import TrainStep
    comm = Communicator.from_env("nccl")
    ds = SyntheticDataset(num_samples, num_nodes=12, num_links=20, num_topologies=3)
    batch = Batch.from_data_list([ds[i] for i in shard_samples(num_samples, rank, world)]).cuda()
    model, graphed = _build(variant)
    step = TrainStep(model, communicator=comm)
    if graphed:
        from gnn_link_prediction_b200.train import GraphedTrainStep
        step_fn = GraphedTrainStep(step)
    else:
        step_fn = step
    losses = [float(step_fn(batch)[0]) for _ in range(3)]
    if rank == 0:
        torch.save({"losses": losses, "params": step.flat_p.cpu()}, os.path.join(out_dir, "dp_gpu.pt"))
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_two_gpu_trainstep_equals_single_gpu(tmp_path, variant):
    """NCCL: 2 ranks on sharded samples == 1 rank on the whole batch, over 3 optimizer steps."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = 6
    mp.spawn(_gpu_worker, args=(2, _free_port(), n, str(tmp_path), variant), nprocs=2, join=True)
    got = torch.load(tmp_path / "dp_gpu.pt")
    from gnn_link_prediction_b200.data import Batch
    from gnn_link_prediction_b200.models import HetroGIN
    from gnn_link_prediction_b200.parallel import Communicator
    from gnn_link_prediction_b200.synthetic import SyntheticDataset
    from gnn_link_prediction_b200.train import TrainStep
    ds = SyntheticDataset(n, num_nodes=12, num_links=20, num_topologies=3)
    batch = Batch.from_data_list([ds[i] for i in range(n)]).cuda()
    model, _ = _build(variant)
    step = TrainStep(model, communicator=Communicator(enabled=False))
    losses = [float(step(batch)[0]) for _ in range(3)]
    torch.testing.assert_close(torch.tensor(got["losses"]), torch.tensor(losses), rtol=1e-5, atol=0)
    if "bn" in variant:
        # the bias of a Linear that feeds a BatchNorm1d has an identically zero gradient: Adam turns its rounding noise
        # into +-lr steps (tests/test_model_gpu.py: final_state_close); such entries may drift apart by 2 * steps * lr
        err = (got["params"] - step.flat_p.cpu()).abs()
        loose = err > 1e-4 * step.flat_p.cpu().abs() + 1e-6
        assert float(loose.float().mean()) < 0.03 and float(err.max()) <= 2.1 * 3 * 1e-3
    else:
        torch.testing.assert_close(got["params"], step.flat_p.cpu(), rtol=1e-4, atol=1e-6)
