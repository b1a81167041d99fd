from . import basesolver, cls_sparsity  # noqa: F401
