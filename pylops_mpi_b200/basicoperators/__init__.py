from .FirstDerivative import MPIFirstDerivative  # noqa: F401
from .BlockDiag import MPIBlockDiag, MPIStackedBlockDiag  # noqa: F401
from .VStack import MPIVStack, MPIStackedVStack  # noqa: F401
from .HStack import MPIHStack  # noqa: F401
from .MatrixMult import (MPIMatrixMult, active_grid_comm, block_gather, local_block_split)  # noqa: F401
from .SecondDerivative import MPISecondDerivative  # noqa: F401
from .Laplacian import MPILaplacian  # noqa: F401
from .Gradient import MPIGradient  # noqa: F401
