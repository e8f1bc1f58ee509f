"""TEST INFRASTRUCTURE (see ../README.md): import-only stand-ins for torch_sparse.

The reference imports `SparseTensor, matmul` (models.py:12) but only uses them in
`GINConv.message_and_aggregate` (models.py:222-225), which is never taken: edge indices
are dense [2,E] tensors (generateFiles.py:176-181), not SparseTensor.
"""


class SparseTensor:  # never instantiated on the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("torch_sparse.SparseTensor is not part of the restated path")


def matmul(*a, **k):
    raise NotImplementedError("torch_sparse.matmul is not part of the restated path")
