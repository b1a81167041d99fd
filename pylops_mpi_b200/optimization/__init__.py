from .basic import cg, cgls  # noqa: F401
from .cls_basic import CG, CGLS  # noqa: F401
