from .._derivatives import FirstDerivative, SecondDerivative  # noqa: F401
from .. import MatrixMult  # noqa: F401
