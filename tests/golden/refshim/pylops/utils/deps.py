def cupy_import(message=None):
    return "cupy not installed (refshim)"
cupy_enabled = False
