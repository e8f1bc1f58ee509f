"""Flat sample container and HBM-resident dataset with on-GPU batch assembly.

The reference keeps one pickled PyG object per sample (`torch.save(data, ...)`,
generateFiles.py:231; `torch.save(torch_data, converted_path)`, dataset.py:118-121), reloads
them with `torch.load` (dataset.py:160-163), collates on the host
(`torch_geometric.loader.DataLoader`, dataset.py:242) and ships every batch over PCIe
(`sample.cuda()`, train.py:28).  Here:

* `SampleArena`  — every field of every sample back to back in ONE flat array per field, plus an
  int64 row-pointer table per size class (path / link / node rows, edges per relation).  It is
  both the in-memory layout and the on-disk format (`save` / `load`: a JSON header followed by
  the raw little-endian arrays, 64-byte aligned, readable with numpy.memmap and nothing else —
  no PyG, no pickle).  Adjacency is stored per sample as the int32 CSRs the kernels consume
  (built once per sample by K0, `data.attach_csr`), optionally with the reference's COO lists.
* `DeviceDataset` — the arena uploaded to HBM once (a B200 holds ~800 k datanet samples in
  180 GB); `collate(ids)` assembles a batch ON THE GPU (`hgin_collate_offsets`,
  `hgin_collate_gather`): a step's host->device traffic is the id list, and no Python loop over
  samples remains on the step's critical path (the host collate of 1024 samples costs ~1 s).
* `DeviceLoader`  — epoch iterator over shuffled ids (`DataLoader(ds, batch_size, shuffle)` of
  dataset.py:242-244), sharded by rank for data parallelism.

Integer work (offsets, CSR concatenation) is bit-exact with `data.Batch.from_data_list(csr=True)`;
tests compare the two.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib
from .data import CONV_EDGE_TYPES, CSR_KEYS, Batch, HeteroData, attach_csr

MAGIC = b"HGINARN1"
NODE_TYPES = ("path", "link", "node")
_NP = {"f32": np.float32, "i32": np.int32, "i64": np.int64}


def _rel_name(et):
    return "__".join(et)


def _as_tensor(a):
    """Zero-copy tensor view of a (possibly read-only, memory-mapped) numpy array; it is only read."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(np.ascontiguousarray(a))


class SampleArena:
    """Host-side flat container.  `arrays[name]` is a contiguous numpy array ("path.x", "path.y",
    "link.x", "node.x", "<src>__<rel>__<dst>.csr_{dst,src}_{rowptr,col}" and optionally
    "....edge_index"), `ptr[cls]` an int64 [S+1] table of first rows per size class ("path", "link",
    "node", "E:<src>__<rel>__<dst>")."""

    def __init__(self, num_samples, edge_types, arrays, ptr, has_coo):
        self.num_samples = int(num_samples)
        self.edge_types = [tuple(et) for et in edge_types]
        self.arrays, self.ptr, self.has_coo = arrays, ptr, bool(has_coo)

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_samples(cls, samples, edge_types=CONV_EDGE_TYPES, keep_coo=True):
        """`samples`: HeteroData objects as the datasets yield them (x / y per node type,
        `edge_index` per relation).  The per-sample CSRs are built by K0 on the GPU
        (`data.attach_csr`; samples sharing a topology tensor build once)."""
        edge_types = [tuple(et) for et in edge_types]
        samples = list(samples)
        if not samples:
            raise ValueError("SampleArena.from_samples: no samples")
        for s in samples:
            attach_csr(s, edge_types)
        ptr, arrays = {}, {}
        for nt in NODE_TYPES:
            counts = np.array([s[nt]["x"].shape[0] for s in samples], dtype=np.int64)
            ptr[nt] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            arrays[f"{nt}.x"] = np.ascontiguousarray(
                torch.cat([s[nt]["x"].to(torch.float32) for s in samples], 0).numpy())
        arrays["path.y"] = np.ascontiguousarray(
            torch.cat([s["path"]["y"].reshape(-1).to(torch.float32) for s in samples], 0).numpy())
        for et in edge_types:
            name = _rel_name(et)
            counts = np.array([s[et]["csr_dst_col"].shape[0] for s in samples], dtype=np.int64)
            ptr["E:" + name] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            for key in CSR_KEYS:
                arrays[f"{name}.{key}"] = np.ascontiguousarray(
                    torch.cat([s[et][key].to(torch.int32) for s in samples], 0).numpy())
            if keep_coo:
                arrays[f"{name}.edge_index"] = np.ascontiguousarray(
                    torch.cat([s[et]["edge_index"].to(torch.int32) for s in samples], 1).numpy())
        return cls(len(samples), edge_types, arrays, ptr, keep_coo)

    # ---- access ---------------------------------------------------------------------------------
    def __len__(self):
        return self.num_samples

    def sizes(self, cls):
        p = self.ptr[cls]
        return p[1:] - p[:-1]

    def rowptr_ptr(self, node_type):
        """First entry of every sample's LOCAL row-pointer block (n + 1 entries per sample)."""
        return self.ptr[node_type] + np.arange(self.num_samples + 1, dtype=np.int64)

    def sample(self, i, index_dtype=torch.int64):
        """Sample `i` as the HeteroData the reference's dataset would return (dataset.py:89-117):
        `edge_index` in `index_dtype` (int64 in the reference) when the arena keeps the COO lists,
        plus the cached per-sample CSRs."""
        if not 0 <= i < self.num_samples:
            raise IndexError(i)
        out = HeteroData()
        for nt in NODE_TYPES:
            lo, hi = self.ptr[nt][i], self.ptr[nt][i + 1]
            out[nt]["x"] = _as_tensor(self.arrays[f"{nt}.x"][lo:hi])
        lo, hi = self.ptr["path"][i], self.ptr["path"][i + 1]
        out["path"]["y"] = _as_tensor(self.arrays["path.y"][lo:hi])
        for et in self.edge_types:
            name = _rel_name(et)
            elo, ehi = self.ptr["E:" + name][i], self.ptr["E:" + name][i + 1]
            if self.has_coo:
                out[et]["edge_index"] = _as_tensor(self.arrays[f"{name}.edge_index"][:, elo:ehi]).to(index_dtype)
            for side, rows_t in (("dst", et[2]), ("src", et[0])):
                rp = self.rowptr_ptr(rows_t)
                out[et][f"csr_{side}_rowptr"] = _as_tensor(self.arrays[f"{name}.csr_{side}_rowptr"][rp[i]:rp[i + 1]])
                out[et][f"csr_{side}_col"] = _as_tensor(self.arrays[f"{name}.csr_{side}_col"][elo:ehi])
        return out

    def __getitem__(self, i):
        return self.sample(i)

    def nbytes(self):
        return sum(a.nbytes for a in self.arrays.values()) + sum(p.nbytes for p in self.ptr.values())

    # ---- collate description (shared by the device and the host collate) ---------------------------
    @property
    def classes(self):
        """Size classes: node types, then edges per relation."""
        return list(NODE_TYPES) + ["E:" + _rel_name(et) for et in self.edge_types]

    def field_specs(self):
        """[(store key, field name, array name, ptr kind, width, size class, add class, closing row, dtype)]:
        how every array is laid out per sample and what offset its int32 entries receive in a batch
        (see hgin_collate_field in include/hgin.h).  ptr kind: a size-class name, or ("rowptr", node type)
        for local row-pointer blocks (n + 1 entries per sample)."""
        cidx = {c: i for i, c in enumerate(self.classes)}
        specs = []
        for nt in NODE_TYPES:
            specs.append((nt, "x", f"{nt}.x", nt, int(self.arrays[f"{nt}.x"].shape[1]), cidx[nt], -1, 0, torch.float32))
        specs.append(("path", "y", "path.y", "path", 1, cidx["path"], -1, 0, torch.float32))
        for et in self.edge_types:
            name = _rel_name(et)
            ec = cidx["E:" + name]
            for side, rows_t, cols_t in (("dst", et[2], et[0]), ("src", et[0], et[2])):
                specs.append((et, f"csr_{side}_rowptr", f"{name}.csr_{side}_rowptr", ("rowptr", rows_t), 1, cidx[rows_t], ec,
                              1, torch.int32))
                specs.append((et, f"csr_{side}_col", f"{name}.csr_{side}_col", "E:" + name, 1, ec, cidx[cols_t], 0,
                              torch.int32))
        return specs

    def _host_tables(self):
        """Contiguous host copies of the ptr tables the C collate reads (built once)."""
        t = self.__dict__.get("_tables")
        if t is None:
            classes = self.classes
            t = {"class_ptr": np.ascontiguousarray(np.stack([np.asarray(self.ptr[c], dtype=np.int64) for c in classes])),
                 "class_sizes": np.stack([self.sizes(c) for c in classes]),
                 "rowptr": {nt: np.ascontiguousarray(self.rowptr_ptr(nt)) for nt in NODE_TYPES}}
            self.__dict__["_tables"] = t
        return t

    def collate_packed(self, ids, pin=True, edge_bucket=None, num_threads=0, out=None):
        """Host collate of samples `ids` straight into ONE packed (optionally pinned) buffer: a
        `data.PackedBatch` whose views equal `Batch.from_data_list([self[i] for i in ids], int32, csr=True,
        keep_coo=False)` plus the `ptr` tables.  Native, multi-threaded (hgin_host_collate); feed it to
        `data.DevicePrefetcher` or `train.GraphedTrainStep`.  `edge_bucket` pads the CSR column arrays to a
        multiple (static shapes for CUDA-graph replay); the tail is never read (row pointers bound it).
        `out`: a uint8 (pinned) tensor to assemble into — reuse a small ring of them in a loader, a fresh
        230 MB allocation costs more than the collate itself."""
        from .data import PackedBatch
        ids_np = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
        B = int(ids_np.shape[0])
        if B == 0:
            raise ValueError("SampleArena.collate_packed: empty id list")
        if ids_np.min() < 0 or ids_np.max() >= self.num_samples:
            raise IndexError(f"sample ids must be in [0, {self.num_samples})")
        tables = self._host_tables()
        classes = self.classes
        C = len(classes)
        totals = tables["class_sizes"][:, ids_np].sum(axis=1)
        specs = self.field_specs()
        layout, off = [], (C * (B + 1) * 8 + 255) // 256 * 256
        places = []
        for key, name, arr_name, ptr_kind, width, sc, ac, closing, dtype in specs:
            rows = int(totals[sc]) + (1 if closing else 0)
            alloc_rows = rows
            if edge_bucket and name.endswith("_col"):
                alloc_rows = (rows + edge_bucket - 1) // edge_bucket * edge_bucket
            shape = (alloc_rows, width) if name == "x" else (alloc_rows,)
            layout.append(("node" if isinstance(key, str) else "edge", key, name, dtype, shape, off))
            places.append((off, rows, alloc_rows))
            off += (alloc_rows * width * 4 + 255) // 256 * 256
        for ci, nt in enumerate(NODE_TYPES):                      # the offsets table doubles as the `ptr` vectors
            layout.append(("node", nt, "ptr", torch.int64, (B + 1,), ci * (B + 1) * 8))
        if out is not None:
            if out.dtype != torch.uint8 or out.is_cuda or out.numel() < off:
                raise ValueError(f"collate_packed: `out` must be a host uint8 tensor of at least {off} bytes")
            buf = out[:off]
        else:
            buf = torch.empty(off, dtype=torch.uint8, pin_memory=bool(pin and torch.cuda.is_available()))
        base = buf.data_ptr()
        table = (_lib.CollateField * len(specs))()
        for i, ((key, name, arr_name, ptr_kind, width, sc, ac, closing, dtype), (o, rows, alloc_rows)) in enumerate(
                zip(specs, places)):
            arr = self.arrays[arr_name]
            ptr = tables["rowptr"][ptr_kind[1]] if isinstance(ptr_kind, tuple) else tables["class_ptr"][classes.index(ptr_kind)]
            table[i] = _lib.CollateField(arr.ctypes.data if arr.size else None, base + o, ptr.ctypes.data, width, sc, ac, closing)
            if alloc_rows > rows:
                buf[o + rows * width * 4:o + alloc_rows * width * 4].zero_()
        _lib.check(_lib.load().hgin_host_collate(B, ids_np.ctypes.data, self.num_samples, len(specs), table, C,
                                                 tables["class_ptr"].ctypes.data, base, int(num_threads)),
                   "hgin_host_collate")
        return PackedBatch(buf, layout, B)

    # ---- on-disk format ---------------------------------------------------------------------------
    def save(self, path):
        """MAGIC | u32 version | u32 header bytes | JSON header | arrays (64-byte aligned)."""
        entries, blobs, off = [], [], 0

        def add(kind, name, arr):
            nonlocal off
            arr = np.ascontiguousarray(arr)
            dtype = {np.dtype(np.float32): "f32", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[arr.dtype]
            entries.append({"kind": kind, "name": name, "dtype": dtype, "shape": list(arr.shape), "offset": off})
            blobs.append(arr)
            off += (arr.nbytes + 63) // 64 * 64

        for name, p in self.ptr.items():
            add("ptr", name, p)
        for name, a in self.arrays.items():
            add("array", name, a)
        header = json.dumps({"num_samples": self.num_samples, "edge_types": [list(et) for et in self.edge_types],
                             "has_coo": self.has_coo, "entries": entries}).encode()
        header += b" " * (-(len(MAGIC) + 8 + len(header)) % 64)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(MAGIC)
            f.write(np.array([1, len(header)], dtype="<u4").tobytes())
            f.write(header)
            for arr in blobs:
                f.write(arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes())
                f.write(b"\0" * (-arr.nbytes % 64))
        os.replace(tmp, path)

    @classmethod
    def load(cls, path, mmap=True):
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path}: not a HeteroGIN sample arena (bad magic)")
            version, hlen = np.frombuffer(f.read(8), dtype="<u4")
            if version != 1:
                raise ValueError(f"{path}: unsupported arena version {version}")
            header = json.loads(f.read(int(hlen)).decode())
        base = len(MAGIC) + 8 + int(hlen)
        size = os.path.getsize(path)
        raw = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
        ptr, arrays = {}, {}
        for e in header["entries"]:
            dt = np.dtype(_NP[e["dtype"]]).newbyteorder("<")
            n = int(np.prod(e["shape"])) if e["shape"] else 1
            lo = base + e["offset"]
            if lo + n * dt.itemsize > size:
                raise ValueError(f"{path}: truncated ({e['name']} ends past the file)")
            arr = raw[lo:lo + n * dt.itemsize].view(dt).reshape(e["shape"])
            (ptr if e["kind"] == "ptr" else arrays)[e["name"]] = arr
        return cls(header["num_samples"], header["edge_types"], arrays, ptr, header["has_coo"])


class DeviceDataset:
    """A `SampleArena` resident in HBM with on-GPU collate."""

    def __init__(self, arena: SampleArena, device=None):
        if not torch.cuda.is_available():
            raise _lib.HginError("DeviceDataset needs a CUDA device (there is no CPU fallback); "
                                 "use SampleArena / data.Batch.from_data_list on the host")
        self.arena = arena
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_samples = arena.num_samples
        self.edge_types = arena.edge_types
        self.classes = arena.classes
        cidx = {c: i for i, c in enumerate(self.classes)}
        self.class_sizes = np.stack([arena.sizes(c) for c in self.classes])            # host: batch totals
        self.class_ptr = self._up(np.stack([np.asarray(arena.ptr[c]) for c in self.classes]))
        rp_ptr = {nt: self._up(arena.rowptr_ptr(nt)) for nt in NODE_TYPES}
        # field table: (store key, field name, device array, ptr table, width, size class, add class, closing row, dtype)
        self.fields = []
        for key, name, arr_name, ptr_kind, width, sc, ac, closing, dtype in arena.field_specs():
            ptr = rp_ptr[ptr_kind[1]] if isinstance(ptr_kind, tuple) else self.class_ptr[cidx[ptr_kind]]
            self.fields.append((key, name, self._up(np.asarray(arena.arrays[arr_name])), ptr, width, sc, ac, closing, dtype))
        widths = np.array([f[4] for f in self.fields], dtype=np.int64)
        per_sample = np.stack([self.class_sizes[f[5]] for f in self.fields]) * widths[:, None]
        self.max_words = int(per_sample.max()) + 1 if per_sample.size else 1

    def _up(self, arr):
        return _as_tensor(arr).to(self.device)

    def __len__(self):
        return self.num_samples

    def nbytes(self):
        return sum(f[2].numel() * f[2].element_size() for f in self.fields)

    def collate(self, ids):
        """Batch of samples `ids` (sequence / numpy / tensor of sample indices), assembled on the GPU.
        Equal to `Batch.from_data_list([ds[i] for i in ids], index_dtype=int32, csr=True,
        keep_coo=False)` moved to the device; `validate()` on the result checks the id range."""
        ids_np = np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids, dtype=np.int64).reshape(-1)
        B = int(ids_np.shape[0])
        if B == 0:
            raise ValueError("DeviceDataset.collate: empty id list")
        if ids_np.min() < 0 or ids_np.max() >= self.num_samples:
            raise IndexError(f"sample ids must be in [0, {self.num_samples})")
        totals = self.class_sizes[:, ids_np].sum(axis=1)                       # host: sizes of the outputs
        ids_host = torch.from_numpy(ids_np.astype(np.int32)).pin_memory()
        ids_dev = ids_host.to(self.device, non_blocking=True)                  # the step's only H2D copy
        C = len(self.classes)
        # one allocation for the whole batch: offsets table, then every field (256-byte aligned)
        layout, off = [], C * (B + 1) * 8
        off = (off + 255) // 256 * 256
        for key, name, src, _, width, sc, _, closing, dtype in self.fields:
            rows = int(totals[sc]) + (1 if closing else 0)
            layout.append((off, rows))
            off += (rows * width * 4 + 255) // 256 * 256
        buf = torch.empty(off, dtype=torch.uint8, device=self.device)
        offsets = buf[:C * (B + 1) * 8].view(torch.int64).view(C, B + 1)
        status = torch.empty(1, dtype=torch.int32, device=self.device)
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.hgin_collate_offsets(B, ids_dev.data_ptr(), C, self.class_ptr.data_ptr(), self.num_samples,
                                            offsets.data_ptr(), status.data_ptr(), stream), "hgin_collate_offsets")
        table = (_lib.CollateField * len(self.fields))()
        out = Batch()
        for i, ((key, name, src, ptr, width, sc, ac, closing, dtype), (o, rows)) in enumerate(zip(self.fields, layout)):
            dst = buf[o:o + rows * width * 4].view(dtype)
            dst = dst.view(rows, width) if name == "x" else dst
            table[i] = _lib.CollateField(src.data_ptr(), dst.data_ptr(), ptr.data_ptr(), width, sc, ac, closing)
            out[key][name] = dst
        _lib.check(lib.hgin_collate_gather(B, ids_dev.data_ptr(), self.num_samples, len(self.fields), table, C,
                                           offsets.data_ptr(), self.max_words, stream), "hgin_collate_gather")
        for ci, nt in enumerate(NODE_TYPES):
            out[nt]["ptr"] = offsets[ci]            # int64 [B+1]: `batch` vectors derive from it lazily
        out.__dict__["num_graphs"] = B
        out.__dict__["_collate_status"] = status
        out.__dict__["_collate_keepalive"] = (ids_host, ids_dev, buf)
        return out

    def h2d_bytes(self, batch_size):
        return 4 * batch_size


def global_steps(num_samples, batch_size, world):
    """Steps per epoch that EVERY rank takes: global chunks of batch_size * world samples; the last partial
    chunk counts only if it gives every rank at least one sample."""
    step = batch_size * world
    full, tail = divmod(num_samples, step)
    return full + (1 if tail >= world else 0)


def shard_id_batches(order, batch_size, rank, world):
    """Rank `rank`'s id list for every global chunk of `order` (positions rank, rank+world, ...).  A final
    chunk with fewer than `world` samples is dropped on all ranks: a rank with an empty shard would skip
    the step and leave the others waiting in the all-reduce."""
    step = batch_size * world
    for lo in range(0, len(order), step):
        chunk = order[lo:lo + step]
        if len(chunk) < world:
            return
        yield chunk[rank::world]


class DeviceLoader:
    """`DataLoader(dataset, batch_size, shuffle)` (dataset.py:242-244) over a `DeviceDataset`:
    yields batches assembled on the GPU; reshuffles every epoch with `generator`; `rank`/`world`
    shard the ids (rank r takes positions r, r+world, ... of every global batch), keeping the last
    partial batch like the reference's loader.  Under data parallelism every rank must enter the two
    all-reduces of a step, so a final chunk with fewer samples than ranks (some shards would be empty)
    is dropped on ALL ranks; `len()` counts the steps every rank takes."""

    def __init__(self, dataset: DeviceDataset, batch_size=1, shuffle=False, generator=None, rank=0, world=1):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator
        self.rank, self.world = rank, world

    def __len__(self):
        return global_steps(len(self.dataset), self.batch_size, self.world)

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator).numpy() if self.shuffle else np.arange(n)
        for ids in shard_id_batches(order, self.batch_size, self.rank, self.world):
            yield self.dataset.collate(ids)


class HostLoader:
    """`DataLoader(dataset, batch_size, shuffle)` (dataset.py:242-244) for a dataset that stays in HOST
    memory (e.g. a memory-mapped `SampleArena` larger than HBM): a background thread assembles packed
    batches with the native collate (`SampleArena.collate_packed`) into a ring of pinned buffers while
    the GPU works; iterate it through `data.DevicePrefetcher` (one DMA per batch):

        for batch in DevicePrefetcher(HostLoader(arena, batch_size=1024, shuffle=True)):
            loss = step(batch)

    A ring slot is reused only after the H2D copy of the batch it held has completed (the prefetcher
    leaves its copy event on the PackedBatch).  `rank` / `world` shard like `DeviceLoader`."""

    def __init__(self, arena: SampleArena, batch_size=1, shuffle=False, generator=None, rank=0, world=1, ring=4,
                 num_threads=0, edge_bucket=None, pin=True, epochs=1):
        if ring < 3:
            raise ValueError("HostLoader: ring must be >= 3 (one batch being filled, one queued, one being copied)")
        self.arena, self.batch_size, self.shuffle, self.generator = arena, int(batch_size), shuffle, generator
        self.rank, self.world, self.ring, self.num_threads = rank, world, ring, num_threads
        self.edge_bucket, self.pin, self.epochs = edge_bucket, pin, epochs
        self._bufs, self._last = [None] * ring, [None] * ring

    def __len__(self):
        return self.epochs * global_steps(len(self.arena), self.batch_size, self.world)

    def _id_batches(self):
        n = len(self.arena)
        for _ in range(self.epochs):           # `epochs` > 1: one uninterrupted stream, reshuffled per epoch
            order = torch.randperm(n, generator=self.generator).numpy() if self.shuffle else np.arange(n)
            yield from shard_id_batches(order, self.batch_size, self.rank, self.world)

    def __iter__(self):
        import queue
        import threading
        q = queue.Queue(maxsize=self.ring - 2)
        stop = threading.Event()

        def produce():
            try:
                for k, ids in enumerate(self._id_batches()):
                    if stop.is_set():
                        return
                    slot = k % self.ring
                    batch = self._produce(slot, ids)
                    while not stop.is_set():
                        try:
                            q.put(batch, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                q.put(None)
            except BaseException as exc:      # surface loader errors in the consumer
                q.put(exc)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
                # control returns here when the consumer asks for the NEXT batch.  A consumer that stages
                # through DevicePrefetcher / GraphedTrainStep has left its copy event on the batch by then; one
                # that read the host buffer itself (CPU use, its own synchronous copy) is simply done with it.
                item.released = True
        finally:
            stop.set()
            for b in self._last:
                if b is not None:
                    b.abandoned = True

    def _produce(self, slot, ids):
        prev = self._last[slot]
        if prev is not None and torch.cuda.is_available():
            import time
            # still queued, or taken from the queue but not staged yet (rare): wait for the hand-off
            while prev.copied is None and not prev.abandoned and not getattr(prev, "released", False):
                time.sleep(0.0002)
            if prev.copied is not None:
                prev.copied.synchronize()                          # its H2D copy has left this buffer
        buf = self._bufs[slot]
        if buf is not None:
            try:
                batch = self.arena.collate_packed(ids, edge_bucket=self.edge_bucket, num_threads=self.num_threads, out=buf)
            except ValueError:                                     # this batch is larger than the slot: grow it
                buf = None
        if buf is None:
            batch = self.arena.collate_packed(ids, pin=self.pin, edge_bucket=self.edge_bucket, num_threads=self.num_threads)
            grown = torch.empty(int(batch.buffer.numel() * 1.125) + 4096, dtype=torch.uint8,
                                pin_memory=bool(self.pin and torch.cuda.is_available()))
            grown[:batch.buffer.numel()].copy_(batch.buffer)
            self._bufs[slot] = grown
            batch.buffer = grown[:batch.buffer.numel()]
        self._last[slot] = batch
        return batch
