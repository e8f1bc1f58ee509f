"""Batch-object API the train step consumes, without PyG.

The reference step (train.py:25-51) touches exactly this surface of a PyG
`HeteroDataBatch`: `.cuda()` (in place, train.py:28), `.x_dict`, `.edge_index_dict`,
`["path"].batch`, `["path"].y`, `["path"].x.shape[0]` (train.py:34,38,50); `load_model` reads
`dataset[0]['link']['x'].shape[1]` (train.py:117-119); the loader is iterated and `len()`-ed
(train.py:24-25).  PyG's collate (`Batch.from_data_list`, reached through
`torch_geometric.loader.DataLoader`, dataset.py:242) concatenates node tensors on dim 0 and
`edge_index` on dim 1 with per-node-type offsets, in sample order — reproduced here bit-exactly
(SURVEY §8(a) A0).

Host-side only: no arithmetic on features happens here.
"""
from __future__ import annotations

import collections

import torch

# Relation order of the reference's HeteroData (dataset.py:112-117).
EDGE_TYPES = (
    ("path", "uses", "link"),
    ("link", "includes", "path"),
    ("link", "connects", "node"),
    ("node", "has", "link"),
    ("path", "is_connected", "node"),
    ("node", "is_used", "path"),
)
# The four relations HetroGIN wires into HeteroConv (models.py:286-290).
CONV_EDGE_TYPES = EDGE_TYPES[:4]


class Store(dict):
    """dict with attribute access (`store.x` == `store['x']`), like a PyG storage.  A batched node
    store that was collated without its `batch` vector derives it from `ptr` on first access."""

    __slots__ = ()

    def __missing__(self, key):
        if key == "batch" and "ptr" in self:
            ptr = dict.__getitem__(self, "ptr")
            counts = ptr[1:] - ptr[:-1]
            vec = torch.repeat_interleave(torch.arange(counts.numel(), dtype=torch.int64, device=ptr.device), counts)
            dict.__setitem__(self, "batch", vec)
            return vec
        raise KeyError(key)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


class HeteroData:
    """Typed graph container: `d['path'].x`, `d['path','uses','link'].edge_index`."""

    def __init__(self):
        self.__dict__["_nodes"] = {}
        self.__dict__["_edges"] = {}

    # -- stores ---------------------------------------------------------------------------
    def __getitem__(self, key) -> Store:
        if isinstance(key, (tuple, list)):
            return self._edges.setdefault(tuple(key), Store())
        return self._nodes.setdefault(key, Store())

    @property
    def node_types(self):
        return list(self._nodes)

    @property
    def edge_types(self):
        return list(self._edges)

    def _gather(self, attr):
        found = {k: s[attr] for k, s in self._nodes.items() if attr in s}
        found.update({k: s[attr] for k, s in self._edges.items() if attr in s})
        return found

    @property
    def x_dict(self):
        return self._gather("x")

    @property
    def edge_index_dict(self):
        return self._gather("edge_index")

    def __getattr__(self, name):
        if name.endswith("_dict"):
            return self._gather(name[: -len("_dict")])
        raise AttributeError(name)

    # -- movement (in place, returns self: `sample.cuda()` at train.py:28) -----------------
    def apply_(self, fn):
        for store in list(self._nodes.values()) + list(self._edges.values()):
            for k, v in store.items():
                if isinstance(v, torch.Tensor):
                    store[k] = fn(v)
        return self

    def to(self, device, non_blocking=False):
        return self.apply_(lambda t: t.to(device, non_blocking=non_blocking))

    def cuda(self, non_blocking=False):
        return self.to("cuda", non_blocking=non_blocking)

    def cpu(self):
        return self.to("cpu")

    def pin_memory(self):
        return self.apply_(lambda t: t if t.is_pinned() else t.pin_memory())

    def nbytes(self):
        total = 0
        for store in list(self._nodes.values()) + list(self._edges.values()):
            for v in store.values():
                if isinstance(v, torch.Tensor):
                    total += v.numel() * v.element_size()
        return total


CSR_KEYS = ("csr_dst_rowptr", "csr_dst_col", "csr_src_rowptr", "csr_src_col")
# id(edge_index tensor) -> (tensor kept alive, per-sample CSR dict): samples that SHARE one edge_index object (the
# synthetic generator reuses a topology's tensors) build their CSRs once.  Bounded LRU: a dataset that loads a fresh
# tensor per __getitem__ (the reference's torch.load per access, dataset.py:160-163) never hits it, and must not
# grow it without limit — keep the CSRs on the samples / in a SampleArena instead (attach_csr stores them there).
_CSR_CACHE = collections.OrderedDict()
_CSR_CACHE_MAX = 256


def attach_csr(sample, edge_types=None):
    """Per-sample adjacency in kernel layout, built ONCE per sample by K0 on the GPU and kept on the
    host next to the COO list: for every relation the destination-sorted CSR (forward aggregation)
    and the source-sorted one (backward gather) as int32 `csr_{dst,src}_{rowptr,col}`.  Batches are
    block-diagonal with contiguous ids, so the batch CSR is the concatenation of these with offsets
    (`Batch.from_data_list(csr=True)`) and no CSR has to be built inside the training step."""
    from . import ops
    for et in (sample.edge_types if edge_types is None else edge_types):
        store = sample[et]
        if all(k in store for k in CSR_KEYS):
            continue
        ei = store["edge_index"]
        hit = _CSR_CACHE.get(id(ei))
        if hit is not None and hit[0] is not ei:      # id() of a dead tensor reused by a new one
            hit = None
        if hit is not None:
            _CSR_CACHE.move_to_end(id(ei))
        if hit is None:
            n_src, n_dst = sample[et[0]]["x"].shape[0], sample[et[2]]["x"].shape[0]
            dev = ei.cuda()
            by_dst = ops.csr_build(dev, n_src, n_dst, by="dst").validate()
            by_src = ops.csr_build(dev, n_src, n_dst, by="src").validate()
            hit = (ei, {"csr_dst_rowptr": by_dst.rowptr.cpu(), "csr_dst_col": by_dst.col.cpu(),
                        "csr_src_rowptr": by_src.rowptr.cpu(), "csr_src_col": by_src.col.cpu()})
            _CSR_CACHE[id(ei)] = hit
            while len(_CSR_CACHE) > _CSR_CACHE_MAX:
                _CSR_CACHE.popitem(last=False)
        store.update(hit[1])
    return sample


class Batch(HeteroData):
    """Block-diagonal concatenation of samples (PyG `Batch.from_data_list` semantics)."""

    @classmethod
    def from_data_list(cls, samples, index_dtype=None, edge_types=None, batch_vector=True, csr=False,
                       keep_coo=True):
        """`index_dtype=None` keeps the samples' dtype (int64 in the reference,
        generateFiles.py:172-181); `torch.int32` narrows on the host so that only 4-byte
        indices cross PCIe.  `edge_types` restricts the relations that are collated (the
        reference ships all six, of which HetroGIN reads four).  `batch_vector=False` leaves the
        per-node graph ids (`store.batch`, int64, only read when GLOBAL_FEATS is on, models.py:347)
        to be derived from `ptr` on first access instead of shipping them every step."""
        out = cls()
        first = samples[0]
        base = {}
        for nt in first.node_types:
            counts = [s[nt]["x"].shape[0] for s in samples]
            starts = [0]
            for c in counts:
                starts.append(starts[-1] + c)
            base[nt] = starts
            store = out[nt]
            for key in first[nt].keys():
                store[key] = torch.cat([s[nt][key] for s in samples], dim=0)
            store["ptr"] = torch.tensor(starts, dtype=torch.int64)
            if batch_vector:
                store["batch"] = torch.repeat_interleave(
                    torch.arange(len(samples), dtype=torch.int64), torch.tensor(counts, dtype=torch.int64))
        for et in (first.edge_types if edge_types is None else edge_types):
            src, _, dst = et
            pieces = []
            for i, s in enumerate(samples):
                ei = s[et]["edge_index"]
                shift = torch.tensor([[base[src][i]], [base[dst][i]]], dtype=ei.dtype)
                pieces.append(ei + shift)
            if keep_coo or not csr:
                ei = torch.cat(pieces, dim=1)
                if index_dtype is not None:
                    ei = ei.to(index_dtype)
                out[et]["edge_index"] = ei
            if csr:
                # batch CSR = per-sample CSRs (attach_csr) concatenated with node / edge offsets
                for s in samples:
                    attach_csr(s, [et])
                e_counts = [s[et]["csr_dst_col"].shape[0] for s in samples]
                e_base = [0]
                for c in e_counts:
                    e_base.append(e_base[-1] + c)
                for side, rows_t, cols_t in (("dst", dst, src), ("src", src, dst)):
                    rp = [s[et][f"csr_{side}_rowptr"][:-1] + e_base[i] for i, s in enumerate(samples)]
                    rp.append(torch.tensor([e_base[-1]], dtype=torch.int32))
                    out[et][f"csr_{side}_rowptr"] = torch.cat(rp).to(torch.int32)
                    out[et][f"csr_{side}_col"] = torch.cat(
                        [s[et][f"csr_{side}_col"] + base[cols_t][i] for i, s in enumerate(samples)]).to(torch.int32)
        out.__dict__["num_graphs"] = len(samples)
        return out

    @property
    def graph(self):
        """The adjacency the model consumes: a prebuilt `functional.GraphCSR` when the batch was
        collated with `csr=True`, otherwise the COO `edge_index_dict` (the model then runs K0)."""
        from .functional import GraphCSR
        ets = [et for et in self.edge_types if "csr_dst_rowptr" in self[et]]
        if not ets:
            return self.edge_index_dict
        num_nodes = {nt: self[nt]["x"].shape[0] for nt in self.node_types}
        blocks = {nt: self[nt]["ptr"] for nt in self.node_types if "ptr" in self[nt]}     # block-diagonal structure
        return GraphCSR.from_prebuilt({et: self[et] for et in ets}, num_nodes, blocks)


class DataLoader:
    """Minimal stand-in for `torch_geometric.loader.DataLoader(ds, batch_size, shuffle)`
    (dataset.py:242-244): single-threaded (`num_workers=0` in the reference), keeps the last
    partial batch, reshuffles every epoch with `generator`."""

    def __init__(self, dataset, batch_size=1, shuffle=False, generator=None, index_dtype=None,
                 edge_types=None, pin_memory=False, batch_vector=True, csr=False, keep_coo=True):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.generator = generator
        self.index_dtype = index_dtype
        self.edge_types = edge_types
        self.pin_memory = pin_memory
        self.batch_vector = batch_vector
        self.csr, self.keep_coo = csr, keep_coo

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator).tolist() if self.shuffle else list(range(n))
        for lo in range(0, n, self.batch_size):
            batch = Batch.from_data_list([self.dataset[i] for i in order[lo:lo + self.batch_size]],
                                         index_dtype=self.index_dtype, edge_types=self.edge_types,
                                         batch_vector=self.batch_vector, csr=self.csr, keep_coo=self.keep_coo)
            yield batch.pin_memory() if self.pin_memory else batch


class DevicePrefetcher:
    """Iterates device-resident batches while the NEXT batch's host->device copy runs on a side
    stream, hidden behind the current step (what `DataLoader(pin_memory=True)` + `non_blocking`
    buys a PyTorch training loop; the reference copies synchronously with `sample.cuda()`,
    train.py:28).  Source batches should be pinned (`DataLoader(..., pin_memory=True)`).

        for batch in DevicePrefetcher(loader):        # batch tensors already live on the GPU
            loss = step(batch)

    A source of `PackedBatch`es (`pack_batch` in the loader) is staged through a ring of `depth`
    preallocated device buffers: ONE DMA per batch and no allocator traffic in the loop.  A slot is
    overwritten only after the step that read it has finished (event recorded on the compute stream
    when the consumer asks for the next batch)."""

    def __init__(self, batches, device=None, depth=3):
        self.source = batches
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._ring, self._free, self._n = [None] * depth, [None] * depth, 0

    def _stage_packed(self, host):
        slot = self._n % self.depth
        self._n += 1
        nbytes = host.buffer.numel()
        if self._ring[slot] is None or self._ring[slot].numel() < nbytes:
            self._ring[slot] = torch.empty(nbytes + nbytes // 8, dtype=torch.uint8, device=self.device)
        dst = self._ring[slot][:nbytes]
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])     # the step that last read this slot is done
            dst.copy_(host.buffer, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        host.copied = ready          # a loader that recycles pinned buffers waits on this before refilling
        return host.views(dst), ready, slot

    def _stage(self, host):
        if host is None:
            return None
        if isinstance(host, PackedBatch):
            return self._stage_packed(host)
        dev = Batch()
        with torch.cuda.stream(self.stream):
            for nt in host.node_types:
                for k, v in host[nt].items():
                    dev[nt][k] = v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v
            for et in host.edge_types:
                for k, v in host[et].items():
                    dev[et][k] = v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v
            ready = torch.cuda.Event()
            ready.record(self.stream)
        dev.__dict__["num_graphs"] = getattr(host, "num_graphs", None)
        return dev, ready, None

    def __iter__(self):
        it = iter(self.source)
        staged = self._stage(next(it, None))
        while staged is not None:
            dev, ready, slot = staged
            staged = self._stage(next(it, None))        # next copy overlaps this batch's step
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ready)
            if slot is None:
                for store in list(dev._nodes.values()) + list(dev._edges.values()):
                    for v in store.values():
                        if isinstance(v, torch.Tensor):
                            v.record_stream(cur)                # allocator: memory is in use on `cur`
            yield dev
            if slot is not None:                                # the consumer has launched its step on `cur`
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self._free[slot] = ev


class PackedBatch:
    """A batch laid out in ONE contiguous byte buffer (every tensor at a 256-byte-aligned offset,
    edge lists padded to `edge_bucket` with (-1, -1) slots that hgin_csr_build drops), so that a
    step's whole input crosses PCIe as a single copy and lands in the static buffers of a captured
    CUDA graph.  `views()` exposes it through the usual Batch API."""

    def __init__(self, buffer, layout, num_graphs):
        self.buffer, self.layout, self.num_graphs = buffer, layout, num_graphs
        self.copied = None        # CUDA event of the H2D copy out of `buffer` (set by DevicePrefetcher)
        self.abandoned = False    # the consumer went away without copying it
        self.released = False     # the consumer asked its loader for the next batch (it is done with this one)

    @property
    def signature(self):
        return tuple((kind, key, name, dtype, shape) for kind, key, name, dtype, shape, _ in self.layout)

    def views(self, buffer=None):
        buf = self.buffer if buffer is None else buffer
        out = Batch()
        for kind, key, name, dtype, shape, off in self.layout:
            n = 1
            for d in shape:
                n *= d
            nbytes = n * torch.empty(0, dtype=dtype).element_size()
            out[key][name] = buf[off:off + nbytes].view(dtype).view(shape)
        out.__dict__["num_graphs"] = self.num_graphs
        return out

    def nbytes(self):
        return self.buffer.numel()


def pack_batch(batch, edge_bucket=8192, pin=True):
    """Batch -> PackedBatch (host side; run it in the loader, next to the collate)."""
    items = []
    for nt in batch.node_types:
        for name, v in batch[nt].items():
            if isinstance(v, torch.Tensor):
                items.append(("node", nt, name, v.contiguous()))
    for et in batch.edge_types:
        for name, v in batch[et].items():
            if not isinstance(v, torch.Tensor):
                continue
            if name == "edge_index":          # COO: pad with (-1,-1) slots, which K0 drops
                e_pad = (v.shape[1] + edge_bucket - 1) // edge_bucket * edge_bucket
                padded = torch.full((2, e_pad), -1, dtype=v.dtype)
                padded[:, :v.shape[1]] = v
                v = padded
            elif name.endswith("_col"):       # prebuilt CSR columns: rowptr bounds them, the tail is never read
                e_pad = (v.shape[0] + edge_bucket - 1) // edge_bucket * edge_bucket
                padded = torch.zeros(e_pad, dtype=v.dtype)
                padded[:v.shape[0]] = v
                v = padded
            items.append(("edge", et, name, v.contiguous()))
    layout, off = [], 0
    for kind, key, name, v in items:
        layout.append((kind, key, name, v.dtype, tuple(v.shape), off))
        off += (v.numel() * v.element_size() + 255) // 256 * 256
    buf = torch.empty(off, dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
    packed = PackedBatch(buf, layout, getattr(batch, "num_graphs", None))
    dst = packed.views()
    for kind, key, name, v in items:
        dst[key][name].copy_(v)
    return packed
