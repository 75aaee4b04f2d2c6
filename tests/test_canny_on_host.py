"""The Canny kernels on the CPU: ``k_canny_nms`` (DP4A Sobel + TG22 non-maximum suppression -> two bit masks),
``k_canny_hyst_smem`` (carry-trick row flood, dirty-row worklist), ``k_canny_expand`` and ``k_hist_finalize`` are compiled from
canny.cu's source text -- launch code (``launch_canny_bits``: grids, strip heights, shared-memory sizes) included -- against
the emulation layer of ``tests/_cuda_emu.py`` and must reproduce ``cv2.Canny(gray, 150, 100, 3, L2gradient=True)`` bit for bit,
and the patch histogram of ``generateDistHistogram`` (imgpyramidrgbd.cpp:146-172).  The `-m gpu` pyramid tests check the same
on the device."""
import ctypes as C

import numpy as np
import pytest

from conftest import synth_pair

# an emulation deadlock must not hang the suite (the C call cannot be interrupted by a signal: kill the run instead)
pytestmark = pytest.mark.timeout(900, method="thread")


@pytest.fixture(scope="module")
def canny_emu(tmp_path_factory):
    import _cuda_emu

    return _cuda_emu.build_canny(str(tmp_path_factory.mktemp("canny_emu")))


def run_canny(lib, gray, t1=150, t2=100, patch=10):
    h, w = gray.shape
    gray = np.ascontiguousarray(gray, np.uint8)
    edges, orig = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
    hist = np.zeros((h // patch, w // patch), np.uint8)
    nz = C.c_int(-1)
    lo, hi = min(t1, t2), max(t1, t2)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    rc = lib.emu_canny_bits(vp(gray), C.c_int(w), C.c_int(h), C.c_int(lo * lo), C.c_int(hi * hi), C.c_int(patch), vp(edges), vp(orig), vp(hist),
                            C.byref(nz))
    assert rc == 0
    return edges, orig, hist, nz.value


@pytest.mark.parametrize("w,h,seed", [(320, 240, 2), (160, 120, 5), (136, 60, 7)])
def test_canny_kernels_on_host_match_opencv(canny_emu, orc32, w, h, seed):
    import cv2

    if (w, h) == (136, 60):          # not a renderer size: noise + blobs, width not a multiple of the 128-pixel warp span
        rng = np.random.default_rng(seed)
        g = cv2.GaussianBlur((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), (0, 0), 2.0)
        gray = cv2.add(g, rng.integers(0, 30, (h, w)).astype(np.uint8))
    else:
        bgr, _ = synth_pair(seed, w, h)["key"]
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    want = cv2.Canny(gray, 150, 100, apertureSize=3, L2gradient=True)
    patch = 10
    edges, orig, hist, nz = run_canny(canny_emu, gray, patch=patch)
    assert want.any() and np.array_equal(edges, want) and np.array_equal(orig, want)
    # generateDistHistogram: edge pixels per patch (u8, wrapping), number of non-empty patches
    cnt = (want[:h // patch * patch, :w // patch * patch] > 0).reshape(h // patch, patch, w // patch, patch).sum(axis=(1, 3))
    assert np.array_equal(hist, (cnt & 255).astype(np.uint8)) and nz == int(((cnt & 255) != 0).sum())


def test_canny_kernels_on_host_edge_cases(canny_emu):
    """Flat image (no gradient anywhere) and a long weak chain hanging on one strong pixel (the hysteresis must follow it
    across band boundaries and in both directions)."""
    import cv2

    flat = np.full((64, 128), 77, np.uint8)
    edges, _, hist, nz = run_canny(canny_emu, flat)
    assert not edges.any() and nz == 0 and not hist.any()
    # a faint diagonal ramp edge (weak everywhere) with one high-contrast spot
    h, w = 120, 160
    img = np.full((h, w), 100, np.uint8)
    for y in range(h):
        img[y, : 20 + y] = 127
    img[60:64, 70:90] = 255
    want = cv2.Canny(img, 150, 100, apertureSize=3, L2gradient=True)
    edges, _, _, _ = run_canny(canny_emu, img)
    assert want.any() and np.array_equal(edges, want)


@pytest.mark.parametrize("w,h", [(1088, 48), (64, 1100)])
def test_canny_kernels_on_host_wide_and_tall_images(canny_emu, w, h):
    """Rows wider than 1024 pixels take the 64-bit-word instantiation of the hysteresis; images with more than 32 rows per warp
    take the global-memory fallback ``k_canny_hyst`` (blind band sweeps).  Same result as cv2.Canny either way."""
    import cv2

    rng = np.random.default_rng(w + h)
    g = cv2.GaussianBlur((rng.integers(0, 2, (h, w)) * 255).astype(np.uint8), (0, 0), 2.5)
    gray = cv2.add(g, rng.integers(0, 25, (h, w)).astype(np.uint8))
    want = cv2.Canny(gray, 150, 100, apertureSize=3, L2gradient=True)
    edges, orig, _, _ = run_canny(canny_emu, gray)
    assert want.any() and np.array_equal(edges, want) and np.array_equal(orig, want)
