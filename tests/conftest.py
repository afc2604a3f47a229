import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # the product library and the oracle are build outputs (git-ignored): build them once if a fresh checkout
    # has neither (nvcc cross-compiles sm_100a without a GPU). The tests never fall back to anything else.
    lib = os.path.join(ROOT, "rttnw_b200", "lib", "librttnw_b200.so")
    orc = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import subprocess
        subprocess.run(["make", "-C", ROOT, "all"], check=True, capture_output=True)


@pytest.fixture(scope="session")
def oracle():
    from tests import _oracle
    return _oracle.load()


@pytest.fixture(scope="session")
def earth_rgba():
    """assets/earth.png (the reference's input fixture, copied byte-for-byte) decoded with PIL."""
    import numpy as np
    from PIL import Image
    path = os.path.join(ROOT, "assets", "earth.png")
    return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8))
