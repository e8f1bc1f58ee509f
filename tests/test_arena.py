"""Flat sample container (on-disk format) and on-GPU batch assembly against the host collate.

CPU part: the container round-trips through its file format and hands back the samples it was
built from (CSR attached by the C oracle so that no GPU is needed).  GPU part (-m gpu): a batch
assembled by `DeviceDataset.collate` equals `Batch.from_data_list(csr=True)` tensor for tensor
(integer work bit-exact), and the train step sees the same numbers from either source."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from gnn_link_prediction_b200.arena import DeviceDataset, DeviceLoader, HostLoader, SampleArena
from gnn_link_prediction_b200.data import CONV_EDGE_TYPES, CSR_KEYS, Batch
from gnn_link_prediction_b200.synthetic import SyntheticDataset


def _attach_csr_with_oracle(sample):
    """What data.attach_csr does with K0 on the GPU, here through the plain-C oracle."""
    for et in CONV_EDGE_TYPES:
        ei = sample[et]["edge_index"].numpy()
        ns, nd = sample[et[0]]["x"].shape[0], sample[et[2]]["x"].shape[0]
        rp, col, _ = c_oracle.csr_build(ei, ns, nd)
        rpt, colt, _ = c_oracle.csr_build(ei[::-1], nd, ns)
        sample[et].update({"csr_dst_rowptr": torch.from_numpy(rp), "csr_dst_col": torch.from_numpy(col),
                           "csr_src_rowptr": torch.from_numpy(rpt), "csr_src_col": torch.from_numpy(colt)})
    return sample


def _ragged_samples(with_csr):
    """Topologies of three different sizes, so every size class is ragged across samples."""
    out = []
    for n, links, seed in ((10, 14, 1), (16, 30, 2), (7, 8, 3)):
        ds = SyntheticDataset(3, num_nodes=n, num_links=links, num_topologies=2, seed=seed)
        out += [ds[i] for i in range(3)]
    return [_attach_csr_with_oracle(s) for s in out] if with_csr else out


def _assert_same_sample(a, b):
    for nt in ("path", "link", "node"):
        assert torch.equal(a[nt]["x"], b[nt]["x"])
    assert torch.equal(a["path"]["y"].reshape(-1), b["path"]["y"].reshape(-1))
    for et in CONV_EDGE_TYPES:
        assert torch.equal(a[et]["edge_index"], b[et]["edge_index"])
        for k in CSR_KEYS:
            assert torch.equal(a[et][k], b[et][k]), (et, k)


def test_arena_round_trips_through_its_file_format(tmp_path):
    samples = _ragged_samples(with_csr=True)
    arena = SampleArena.from_samples(samples)
    assert len(arena) == len(samples)
    for i, s in enumerate(samples):
        _assert_same_sample(arena[i], s)
        assert arena[i][CONV_EDGE_TYPES[0]]["edge_index"].dtype == torch.int64      # generateFiles.py:172-181
    path = str(tmp_path / "train.hgin")
    arena.save(path)
    for mmap in (True, False):
        back = SampleArena.load(path, mmap=mmap)
        assert back.num_samples == arena.num_samples and back.edge_types == arena.edge_types
        assert sorted(back.arrays) == sorted(arena.arrays) and sorted(back.ptr) == sorted(arena.ptr)
        for k in arena.arrays:
            assert back.arrays[k].dtype == arena.arrays[k].dtype and np.array_equal(back.arrays[k], arena.arrays[k]), k
        for k in arena.ptr:
            assert np.array_equal(back.ptr[k], arena.ptr[k]), k
        for i in (0, 4, len(samples) - 1):
            _assert_same_sample(back[i], samples[i])
    with pytest.raises(IndexError):
        arena[len(samples)]


def test_arena_rejects_foreign_and_truncated_files(tmp_path):
    bad = tmp_path / "bad.hgin"
    bad.write_bytes(b"not an arena at all, just bytes" * 10)
    with pytest.raises(ValueError, match="bad magic"):
        SampleArena.load(str(bad))
    arena = SampleArena.from_samples(_ragged_samples(with_csr=True)[:2], keep_coo=False)
    good = tmp_path / "good.hgin"
    arena.save(str(good))
    data = good.read_bytes()
    cut = tmp_path / "cut.hgin"
    cut.write_bytes(data[:len(data) // 2])
    with pytest.raises(ValueError, match="truncated"):
        SampleArena.load(str(cut))
    assert not SampleArena.load(str(good)).has_coo
    with pytest.raises(ValueError):
        SampleArena.from_samples([])


def test_device_dataset_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    arena = SampleArena.from_samples(_ragged_samples(with_csr=True)[:2])
    with pytest.raises(Exception, match="no CPU fallback"):
        DeviceDataset(arena)


def test_device_loader_shards_every_global_batch_across_ranks():
    """Host logic of DeviceLoader (no GPU): rank r takes positions r, r+W, ... of every global batch, all
    ranks walk the same shuffled order, the last partial batch is kept, nothing is seen twice."""

    class Stub:                      # what DeviceLoader needs from a dataset: len() and collate(ids)
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def collate(self, ids):
            return list(int(i) for i in ids)

    # (37, 4, 3): tail of 1 < 3 ranks; (33, 4, 4): tail of 1 < 4 ranks; (5, 8, 4): one chunk of 5 >= 4 ranks is kept
    for n, bs, world in [(37, 4, 3), (64, 8, 2), (5, 8, 4), (16, 4, 1), (33, 4, 4), (3, 8, 4)]:
        per_rank = []
        for rank in range(world):
            loader = DeviceLoader(Stub(n), batch_size=bs, shuffle=True, generator=torch.Generator().manual_seed(9),
                                  rank=rank, world=world)
            batches = list(loader)
            # EVERY rank takes exactly len(loader) steps with a non-empty shard (both all-reduces of a step are
            # entered by all ranks): a final chunk with fewer samples than ranks is dropped everywhere
            assert len(batches) == len(loader) and all(1 <= len(b) <= bs for b in batches)
            per_rank.append(batches)
        step = bs * world
        kept = n - (n % step if n % step < world else 0)
        order = torch.randperm(n, generator=torch.Generator().manual_seed(9)).tolist()
        seen = sorted(i for batches in per_rank for b in batches for i in b)
        assert seen == sorted(order[:kept])                               # every kept sample exactly once
        for g, lo in enumerate(range(0, kept, step)):                    # global batch g, split round-robin
            chunk = order[lo:lo + step]
            for rank in range(world):
                assert per_rank[rank][g] == chunk[rank::world]
    plain = list(DeviceLoader(Stub(10), batch_size=4))
    assert plain == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]                  # shuffle=False keeps dataset order


def _assert_views_equal_host_batch(views, want, num_graphs):
    assert views.num_graphs == num_graphs
    for nt in ("path", "link", "node"):
        assert torch.equal(views[nt]["x"].cpu(), want[nt]["x"])
        assert torch.equal(views[nt]["ptr"].cpu(), want[nt]["ptr"])
    assert torch.equal(views["path"]["y"].cpu(), want["path"]["y"].reshape(-1))
    for et in CONV_EDGE_TYPES:
        for k in CSR_KEYS:
            w = want[et][k]
            assert views[et][k].dtype == torch.int32
            assert torch.equal(views[et][k][:w.shape[0]].cpu(), w), (et, k)     # the tail of a padded col array is never read


@pytest.mark.parametrize("ids", [[0, 1, 2, 3, 4, 5, 6, 7, 8], [8, 3, 3, 0, 7], [5], list(range(9)) * 5])
@pytest.mark.parametrize("edge_bucket", [None, 64])
def test_native_host_collate_equals_python_collate(ids, edge_bucket):
    """hgin_host_collate (C++, threads) into one packed buffer == Batch.from_data_list(csr=True), no GPU needed."""
    samples = _ragged_samples(with_csr=True)
    arena = SampleArena.from_samples(samples, keep_coo=False)
    for threads in (1, 3):
        packed = arena.collate_packed(ids, pin=False, edge_bucket=edge_bucket, num_threads=threads)
        _assert_views_equal_host_batch(packed.views(), _host_batch(samples, ids), len(ids))
    out = torch.empty(packed.nbytes() + 1000, dtype=torch.uint8)
    again = arena.collate_packed(ids, edge_bucket=edge_bucket, out=out)
    assert again.buffer.data_ptr() == out.data_ptr()              # assembled in place (the gaps between fields are padding)
    _assert_views_equal_host_batch(again.views(), _host_batch(samples, ids), len(ids))
    with pytest.raises(ValueError):
        arena.collate_packed(ids, out=torch.empty(16, dtype=torch.uint8))
    with pytest.raises(IndexError):
        arena.collate_packed([0, 99])


def test_host_loader_streams_an_epoch_through_a_buffer_ring():
    samples = _ragged_samples(with_csr=True)
    arena = SampleArena.from_samples(samples, keep_coo=False)
    for epoch in range(2):                                   # the ring is reused across epochs
        loader = HostLoader(arena, batch_size=2, shuffle=False, ring=3, pin=False)
        seen = 0
        for k, packed in enumerate(loader):                  # 5 batches through 3 slots
            ids = list(range(2 * k, min(2 * k + 2, 9)))
            _assert_views_equal_host_batch(packed.views(), _host_batch(samples, ids), len(ids))
            seen += packed.num_graphs
        assert seen == 9 and len(loader) == 5
    per_rank = [[p.num_graphs for p in HostLoader(arena, batch_size=2, shuffle=True, pin=False,
                                                   generator=torch.Generator().manual_seed(1), rank=r, world=2)]
                for r in range(2)]
    # 9 samples, global chunks of 4: two full chunks, then a tail of 1 < 2 ranks that every rank drops
    assert per_rank == [[2, 2], [2, 2]]
    with pytest.raises(ValueError):
        HostLoader(arena, ring=2)


@pytest.mark.gpu
@pytest.mark.timeout(120)
def test_host_loader_iterated_directly_on_a_gpu_box_does_not_wait_for_a_staging_event():
    """A consumer that reads the pinned batches itself (no DevicePrefetcher, so nobody sets `copied`) must not
    stall the producer when the ring wraps: asking for the next batch releases the previous one."""
    samples = _ragged_samples(with_csr=True)
    arena = SampleArena.from_samples(samples, keep_coo=False)
    loader = HostLoader(arena, batch_size=1, shuffle=False, ring=3, pin=True)
    seen = []
    for k, packed in enumerate(loader):                      # 9 batches through 3 pinned slots
        dev = packed.views(packed.buffer.cuda())             # the consumer's own synchronous copy
        _assert_views_equal_host_batch(dev, _host_batch(samples, [k]), 1)
        seen.append(packed.num_graphs)
    assert seen == [1] * 9


# ---- on-GPU collate -----------------------------------------------------------------------------
def _host_batch(samples, ids):
    return Batch.from_data_list([samples[i] for i in ids], index_dtype=torch.int32, edge_types=CONV_EDGE_TYPES,
                                batch_vector=True, csr=True, keep_coo=False)


@pytest.mark.gpu
@pytest.mark.parametrize("ids", [[0, 1, 2, 3, 4, 5, 6, 7, 8], [8, 3, 3, 0, 7], [5], [2, 6], list(range(9)) * 7])
def test_device_collate_equals_host_collate(ids):
    samples = _ragged_samples(with_csr=False)
    ds = DeviceDataset(SampleArena.from_samples(samples))        # CSRs by K0 on the GPU
    got = ds.collate(ids)
    want = _host_batch(samples, ids)
    assert got.num_graphs == len(ids)
    for nt in ("path", "link", "node"):
        assert torch.equal(got[nt]["x"].cpu(), want[nt]["x"])
        assert torch.equal(got[nt]["ptr"].cpu(), want[nt]["ptr"])
        assert torch.equal(got[nt]["batch"].cpu(), want[nt]["batch"])          # derived lazily from ptr
    assert torch.equal(got["path"]["y"].cpu(), want["path"]["y"].reshape(-1))
    for et in CONV_EDGE_TYPES:
        for k in CSR_KEYS:
            assert got[et][k].dtype == torch.int32
            assert torch.equal(got[et][k].cpu(), want[et][k]), (et, k)


@pytest.mark.gpu
def test_device_collate_large_batch_and_loader_sharding():
    ds0 = SyntheticDataset(64, num_topologies=4, seed=11)
    samples = [ds0[i] for i in range(64)]
    ds = DeviceDataset(SampleArena.from_samples(samples, keep_coo=False))
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 64, (300,), generator=g).tolist()
    got, want = ds.collate(ids), _host_batch(samples, ids)
    for et in CONV_EDGE_TYPES:
        for k in CSR_KEYS:
            assert torch.equal(got[et][k].cpu(), want[et][k]), (et, k)
    assert torch.equal(got["path"]["x"].cpu(), want["path"]["x"])
    with pytest.raises(IndexError):
        ds.collate([0, 64])
    with pytest.raises(ValueError):
        ds.collate([])
    # two ranks see disjoint halves of every global batch; together they cover the epoch once
    seen = []
    for rank in range(2):
        loader = DeviceLoader(ds, batch_size=8, shuffle=True, generator=torch.Generator().manual_seed(3), rank=rank, world=2)
        assert len(loader) == 4
        seen.append(sum(b.num_graphs for b in loader))
    assert sum(seen) == 64 and seen[0] == seen[1]


@pytest.mark.gpu
def test_host_loader_through_device_prefetcher_matches_device_collate():
    from gnn_link_prediction_b200.data import DevicePrefetcher
    ds0 = SyntheticDataset(22, num_nodes=12, num_links=20, num_topologies=3, seed=8)
    samples = [ds0[i] for i in range(22)]
    arena = SampleArena.from_samples(samples, keep_coo=False)
    dev = DeviceDataset(arena)
    loader = HostLoader(arena, batch_size=3, shuffle=True, generator=torch.Generator().manual_seed(4), ring=3)
    order = torch.randperm(22, generator=torch.Generator().manual_seed(4)).tolist()
    n = 0
    for k, batch in enumerate(DevicePrefetcher(loader, depth=2)):      # 8 batches: pinned ring and device ring both wrap
        ids = order[3 * k:3 * k + 3]
        want = dev.collate(ids)
        for nt in ("path", "link", "node"):
            assert torch.equal(batch[nt]["x"], want[nt]["x"]) and torch.equal(batch[nt]["ptr"], want[nt]["ptr"])
        for et in CONV_EDGE_TYPES:
            for key in CSR_KEYS:
                assert torch.equal(batch[et][key][:want[et][key].shape[0]], want[et][key]), (et, key)
        n += batch.num_graphs
    assert n == 22


@pytest.mark.gpu
def test_train_step_is_identical_on_device_collated_batches():
    from gnn_link_prediction_b200.models import HetroGIN
    from gnn_link_prediction_b200.train import TrainStep
    ds0 = SyntheticDataset(6, num_nodes=12, num_links=20, num_topologies=3, seed=4)
    samples = [ds0[i] for i in range(6)]
    dev_ds = DeviceDataset(SampleArena.from_samples(samples))
    kw = dict(node_embedding_size=16, message_passing_layers=2, dropout=0.0, concat_path=True, bl_features=False,
              divided_features=False, global_feats=False, mlp_layers=[32, 16], act="torch.nn.PReLU()",
              mlp_head_act=None, mlp_bn=False)
    losses = []
    for source in ("host", "device"):
        torch.manual_seed(0)
        step = TrainStep(HetroGIN(input_channels={"link": 7, "path": 7, "node": 3}, **kw).cuda().train())
        run = []
        for ids in ([0, 1, 2], [3, 4, 5], [5, 0, 3]):
            batch = _host_batch(samples, ids).cuda() if source == "host" else dev_ds.collate(ids)
            run.append(step(batch).clone())
        losses.append(torch.stack(run).cpu())
    assert torch.equal(losses[0], losses[1])
