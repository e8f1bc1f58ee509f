"""Restatement of torch_geometric.loader.DataLoader (dataset.py:242-244): a torch DataLoader
whose collate_fn is Batch.from_data_list."""
import torch
from ..data import Batch


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)
