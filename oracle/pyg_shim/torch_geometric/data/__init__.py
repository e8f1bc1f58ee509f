"""Restatement of the torch_geometric.data subset the reference touches:
`Data.from_dict` (generateFiles.py:187), `HeteroData` stores (dataset.py:89-117),
`Dataset` base class (dataset.py:28) and hetero `Batch.from_data_list` (via DataLoader,
dataset.py:242)."""
import copy
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_dict(cls, mapping):
        return cls(**mapping)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]


class _Storage(dict):
    """Attribute-style dict: `store.x`, `store['x']`."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]


class HeteroData:
    def __init__(self):
        object.__setattr__(self, "_node_stores", {})
        object.__setattr__(self, "_edge_stores", {})

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self._edge_stores.setdefault(tuple(key), _Storage())
        return self._node_stores.setdefault(key, _Storage())

    @property
    def node_types(self):
        return list(self._node_stores.keys())

    @property
    def edge_types(self):
        return list(self._edge_stores.keys())

    def _collect(self, attr):
        out = {}
        for k, s in self._node_stores.items():
            if attr in s:
                out[k] = s[attr]
        for k, s in self._edge_stores.items():
            if attr in s:
                out[k] = s[attr]
        return out

    def __getattr__(self, name):
        if name.endswith("_dict"):
            return self._collect(name[:-5])
        raise AttributeError(name)

    def _apply(self, fn):
        for s in list(self._node_stores.values()) + list(self._edge_stores.values()):
            for k, v in s.items():
                if isinstance(v, torch.Tensor):
                    s[k] = fn(v)
        return self

    def to(self, device):
        return self._apply(lambda t: t.to(device))

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def clone(self):
        return copy.deepcopy(self)


class Batch(HeteroData):
    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        node_types = data_list[0].node_types
        edge_types = data_list[0].edge_types
        offsets = {t: [0] for t in node_types}
        for t in node_types:
            store = out[t]
            keys = list(data_list[0][t].keys())
            for k in keys:
                store[k] = torch.cat([d[t][k] for d in data_list], dim=0)
            n_per = [d[t]["x"].size(0) for d in data_list]
            for n in n_per:
                offsets[t].append(offsets[t][-1] + n)
            store["batch"] = torch.repeat_interleave(torch.arange(len(data_list)), torch.tensor(n_per))
            store["ptr"] = torch.tensor(offsets[t])
        for et in edge_types:
            src, _, dst = et
            parts = []
            for i, d in enumerate(data_list):
                ei = d[et]["edge_index"]
                inc = torch.tensor([[offsets[src][i]], [offsets[dst][i]]], dtype=ei.dtype)
                parts.append(ei + inc)
            out[et]["edge_index"] = torch.cat(parts, dim=1)
        object.__setattr__(out, "num_graphs", len(data_list))
        return out


class Dataset(torch.utils.data.Dataset):
    pass
