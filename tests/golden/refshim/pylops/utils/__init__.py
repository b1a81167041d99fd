from typing import Any
DTypeLike = Any
NDArray = Any
ShapeLike = Any
InputDimsLike = Any
from . import deps, backend, typing, _internal  # noqa: E402,F401
