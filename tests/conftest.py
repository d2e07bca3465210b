import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def small_sets():
    return dict(np.load(os.path.join(GOLDEN_DIR, "small_sets.npz")))


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.ref_available():
        pytest.skip("reference driver oracle/_ref/libvkhr_ref.so not available")
    return oracle.ref()


@pytest.fixture(scope="session")
def vox():
    import vkhr_b200
    v = vkhr_b200.Voxelizer(0)
    yield v
    v.close()


def dense(sparse: dict, n: int) -> np.ndarray:
    d = np.zeros(n, dtype=np.uint8)
    for k, v in sparse.items():
        d[int(k)] = v
    return d
