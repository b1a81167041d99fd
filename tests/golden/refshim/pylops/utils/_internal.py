from numbers import Integral
import numpy as np


def _value_or_sized_to_tuple(value, repeat=1):
    if isinstance(value, (Integral, np.integer)):
        return tuple([int(value)] * repeat)
    return tuple(int(v) for v in value)
