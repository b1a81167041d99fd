from typing import Any
DTypeLike = Any
NDArray = Any
ShapeLike = Any
InputDimsLike = Any
