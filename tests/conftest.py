import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu` through gpurun")
    # the product library and the C oracle are build artefacts; make sure they exist before collection
    # build.py is loaded by path: importing the package itself needs the built library (no CPU fallback)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mxd_build", os.path.join(ROOT, "mxdetection_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_library()
    from oracle import cref
    cref.build()


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
