"""``@reshaped`` (pylops_mpi/utils/decorators.py:9-80) re-expressed for cached
partitions: bring the flat SCATTER input to the operator's row-block (or
stacking) partition -- a no-op view when the partitions already agree, one
grouped NCCL exchange otherwise -- call the body on a ``dims``-shaped array,
ravel the result."""
from __future__ import annotations

from functools import wraps
from typing import Callable, Optional

import numpy as np

from ..DistributedArray import DistributedArray, Partition
from .partition import local_split_sizes


def reshaped(func: Optional[Callable] = None, forward: Optional[bool] = None,
             stacking: Optional[bool] = None) -> Callable:
    def decorator(f):
        @wraps(f)
        def wrapper(self, x: DistributedArray):
            if x.partition is not Partition.SCATTER:
                raise ValueError(f"x should have partition={Partition.SCATTER}, "
                                 f"{x.partition} != {Partition.SCATTER}")
            if stacking and forward:
                local_shapes = getattr(self, "local_shapes_m")
                global_shape = x.global_shape
            elif stacking and not forward:
                local_shapes = getattr(self, "local_shapes_n")
                global_shape = x.global_shape
            else:
                dims = tuple(getattr(self, "dims"))
                global_shape = dims
                ext = local_split_sizes(dims[0], x.size)
                local_shapes = [(e,) + dims[1:] for e in ext]
            dst = [int(np.prod(s)) for s in local_shapes]
            buf = x._repartition_flat(dst).view(local_shapes[x.rank])
            arr = DistributedArray(global_shape=global_shape, base_comm=x.base_comm,
                                   local_shapes=local_shapes, axis=0, dtype=x._tdtype, _buffer=buf,
                                   _trusted=(local_shapes is not None))
            y: DistributedArray = f(self, arr)
            if len(y.global_shape) > 1:
                y = y._ravel_view()      # y is a fresh temporary: flatten without the copy of :74-75
            return y
        return wrapper
    if func is not None:
        return decorator(func)
    return decorator
