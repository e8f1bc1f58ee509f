"""TEST INFRASTRUCTURE (see ../README.md): minimal restatement of torch_geometric 2.0.2."""
from . import typing, utils, data, loader, nn  # noqa: F401

__version__ = "2.0.2-shim"
