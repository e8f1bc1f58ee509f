from . import conv  # noqa: F401
from .conv import MessagePassing, HeteroConv, GATConv  # noqa: F401
from .pool import global_mean_pool, global_max_pool  # noqa: F401
