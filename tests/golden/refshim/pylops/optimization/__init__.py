from . import basesolver  # noqa: F401
