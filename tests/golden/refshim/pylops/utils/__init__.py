from typing import Any
DTypeLike = Any
NDArray = Any
ShapeLike = Any
InputDimsLike = Any
from . import deps, backend, typing, _internal  # noqa: E402,F401
from .backend import get_array_module, get_module_name  # noqa: E402,F401


def get_real_dtype(dtype):
    import numpy as np
    return np.real(np.ones(1, dtype)).dtype
