import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # make sure libb200lops.so exists / is current before any test imports the package
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b200_build", os.path.join(ROOT, "pylops_mpi_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build()
    except Exception as exc:  # e.g. no nvcc on the box: the prebuilt in-tree .so is used
        if not os.path.exists(mod.OUT):
            raise
        print(f"[conftest] using prebuilt {mod.OUT} ({exc})")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
