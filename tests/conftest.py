import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    if os.environ.get("REVO_ASSUME_GPU") == "1":      # numpy/ctypes-only runs on a GPU box: skip the torch import
        return True
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def orc32():
    from oracle import oracle as O

    return O.Oracle("f32")


@pytest.fixture(scope="session")
def orc64():
    from oracle import oracle as O

    return O.Oracle("f64")


@pytest.fixture(scope="session")
def ctx():
    from revo_b200 import api

    c = api.Context(0)
    yield c
    c.close()


_PAIR_CACHE = {}


def synth_pair(seed, w=640, h=480, xi=None):
    from revo_b200 import synth

    key = (seed, w, h, None if xi is None else tuple(np.asarray(xi).tolist()))
    if key not in _PAIR_CACHE:
        _PAIR_CACHE[key] = synth.make_pair(seed, w, h, xi=xi)
    return _PAIR_CACHE[key]


def rot_angle(Ra, Rb):
    """Angle of Ra^T Rb, from the skew part (well conditioned for small angles, unlike arccos(trace))."""
    R = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    s = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = 0.5 * (np.trace(R) - 1.0)
    return float(np.arctan2(s, c))
