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
        # size classes: node types, then edges per relation
        self.classes = list(NODE_TYPES) + ["E:" + _rel_name(et) for et in self.edge_types]
        cidx = {c: i for i, c in enumerate(self.classes)}
        self.class_sizes = np.stack([arena.sizes(c) for c in self.classes])            # host: batch totals
        self.class_ptr = self._up(np.stack([np.asarray(arena.ptr[c]) for c in self.classes]))
        up = lambda name: self._up(np.asarray(arena.arrays[name]))
        # field table: (store key, field name, device array, ptr table, width, size class, add class, closing row)
        self.fields = []
        node_ptr = {nt: self.class_ptr[cidx[nt]] for nt in NODE_TYPES}
        rp_ptr = {nt: self._up(arena.rowptr_ptr(nt)) for nt in NODE_TYPES}
        for nt in NODE_TYPES:
            a = up(f"{nt}.x")
            self.fields.append((nt, "x", a, node_ptr[nt], a.shape[1], cidx[nt], -1, 0, torch.float32))
        self.fields.append(("path", "y", up("path.y"), node_ptr["path"], 1, cidx["path"], -1, 0, torch.float32))
        for et in self.edge_types:
            name = _rel_name(et)
            ec = cidx["E:" + name]
            e_ptr = self.class_ptr[ec]
            for side, rows_t, cols_t in (("dst", et[2], et[0]), ("src", et[0], et[2])):
                self.fields.append((et, f"csr_{side}_rowptr", up(f"{name}.csr_{side}_rowptr"), rp_ptr[rows_t], 1,
                                    cidx[rows_t], ec, 1, torch.int32))
                self.fields.append((et, f"csr_{side}_col", up(f"{name}.csr_{side}_col"), e_ptr, 1, ec, cidx[cols_t], 0,
                                    torch.int32))
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


class DeviceLoader:
    """`DataLoader(dataset, batch_size, shuffle)` (dataset.py:242-244) over a `DeviceDataset`:
    yields batches assembled on the GPU; reshuffles every epoch with `generator`; `rank`/`world`
    shard the ids (rank r takes positions r, r+world, ... of every global batch), keeping the last
    partial batch like the reference's loader."""

    def __init__(self, dataset: DeviceDataset, batch_size=1, shuffle=False, generator=None, rank=0, world=1):
        self.dataset, self.batch_size, self.shuffle, self.generator = dataset, int(batch_size), shuffle, generator
        self.rank, self.world = rank, world

    def __len__(self):
        return (len(self.dataset) + self.batch_size * self.world - 1) // (self.batch_size * self.world)

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator).numpy() if self.shuffle else np.arange(n)
        step = self.batch_size * self.world
        for lo in range(0, n, step):
            ids = order[lo:lo + step][self.rank::self.world]
            if len(ids):
                yield self.dataset.collate(ids)
