import numpy as np


def get_module(backend="numpy"):
    assert backend == "numpy"
    return np


def get_array_module(x):
    return np


def get_module_name(mod):
    return "numpy"


def to_numpy(x):
    return x


def get_normalize_axis_index():
    from numpy.lib.array_utils import normalize_axis_index
    return normalize_axis_index
