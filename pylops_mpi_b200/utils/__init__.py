from .dottest import dottest  # noqa: F401
