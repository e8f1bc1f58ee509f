from .message_passing import MessagePassing  # noqa: F401
from .hetero_conv import HeteroConv  # noqa: F401
from .gat_conv import GATConv  # noqa: F401
