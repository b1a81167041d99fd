from collections.abc import Sized


def _value_or_sized_to_tuple(value, repeat=1):
    if isinstance(value, tuple):
        return value
    if not isinstance(value, Sized):
        return tuple([value] * repeat)
    return tuple(value)
