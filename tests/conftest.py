import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.bindings import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libcanny_ref.so not built (reference sources absent)")
    return Ref()


@pytest.fixture(scope="session")
def test_gray():
    """tests/test.jpg of the reference decoded to gray (256x256), committed as raw bytes so the
    input does not depend on the JPEG decoder (sha256 pinned in tests/golden/manifest.json)."""
    p = ROOT / "tests" / "golden" / "test_gray_256x256.u8"
    return np.fromfile(p, dtype=np.uint8).reshape(256, 256)


@pytest.fixture(scope="session")
def gpu_ctx():
    import canny_edge_b200 as cb
    ctx = cb.Context(0)
    yield ctx
    ctx.close()
