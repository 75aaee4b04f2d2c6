"""Regenerates the committed golden vectors under tests/golden/ (run in the AUTHORING container only: it
reads /root/reference and needs cv2; the GPU box never runs this).

  se3_kat.npz   -- known answers of Sophus::SE3::exp / group product, produced by the reference's OWN sympy
                   implementation (thirdparty/Sophus/py/sophus/se3.py) for the tangent vectors of
                   thirdparty/Sophus/test/core/test_se3.cpp:30-44 and the rotX/rotY/rotZ identities of
                   test_se3.cpp:137-146, plus the motion of BASELINE.md config 1.
  cv2_small.npz -- outputs of the OpenCV kernels the reference calls (imgpyramidrgbd.cpp:53,82,184,241) from
                   python cv2 4.13 on small seeded images: cvtColor, pyrDown, Canny(150,100,3,L2), and
                   distanceTransform(L2, PRECISE) (OpenCV's own path, IPP branch off -- see oracle.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def make_se3():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/thirdparty/Sophus/py")
    import sympy
    from sophus.matrix import Vector6
    from sophus.se3 import Se3

    tangents = [
        [0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 1, 0, 0], [0, -5, 10, 0, 0, 0], [-1, 1, 0, 0, 0, 1],
        [20, -1, 0, -1, 1, 0], [30, 5, -1, 20, -1, 0],            # test_se3.cpp:30-44
        [0, 0, 0, 0.2, 0, 0], [0, 0, 0, 0, -0.2, 0], [0, 0, 0, 0, 0, 1.1],   # test_se3.cpp:137-146 (rotX/rotY/rotZ)
        [0.010, -0.006, 0.008, 0.004, -0.006, 0.003],
        [1e-7, 2e-7, -1e-7, 1e-6, -2e-6, 3e-6],                   # below Sophus' float epsilon (Taylor branch)
    ]
    mats = []
    for v in tangents:
        if all(x == 0 for x in v[3:]):
            # sympy's exp divides by theta; the zero-rotation answer is the pure translation (se3.hpp:735-737)
            M = np.eye(4)
            M[:3, 3] = v[:3]
        else:
            T = Se3.exp(Vector6(*[sympy.Float(x, 40) for x in v]))
            M = np.array(sympy.N(T.matrix(), 30).tolist(), dtype=np.float64)
        mats.append(M)
    mats = np.array(mats)
    # group products exp(a) * exp(b) for consecutive pairs (se3.hpp:285-289)
    prods = np.array([mats[i] @ mats[i + 1] for i in range(len(mats) - 1)])
    # closed-form rotX/rotY/rotZ matrices (the other side of the test_se3.cpp:137-146 identities)
    c, s = np.cos, np.sin
    rot = np.array([
        [[1, 0, 0], [0, c(0.2), -s(0.2)], [0, s(0.2), c(0.2)]],
        [[c(-0.2), 0, s(-0.2)], [0, 1, 0], [-s(-0.2), 0, c(-0.2)]],
        [[c(1.1), -s(1.1), 0], [s(1.1), c(1.1), 0], [0, 0, 1]],
    ])
    np.savez(os.path.join(HERE, "se3_kat.npz"), tangents=np.array(tangents, np.float64), exp=mats, prod=prods, rotxyz=rot)


def make_cv2():
    import cv2

    rng = np.random.default_rng(42)
    out = {}
    # a structured image (rectangles + noise) and a pure-noise image, 96x64
    h, w = 64, 96
    img = np.full((h, w, 3), 60, np.uint8)
    for k in range(7):
        x0, y0 = int(rng.integers(0, w - 20)), int(rng.integers(0, h - 16))
        img[y0:y0 + int(rng.integers(8, 30)), x0:x0 + int(rng.integers(8, 40))] = rng.integers(40, 250, 3)
    img = np.clip(img.astype(np.int16) + rng.integers(-3, 4, img.shape), 0, 255).astype(np.uint8)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    cv2.ipp.setUseIPP(False)
    for name, bgr in (("rect", img), ("noise", noise)):
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        down = cv2.pyrDown(gray)
        edges = cv2.Canny(gray, 150, 100, apertureSize=3, L2gradient=True)
        edges2 = cv2.Canny(gray, 60, 20, apertureSize=3, L2gradient=True)
        dt = cv2.distanceTransform(255 - edges, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
        out.update({f"{name}_bgr": bgr, f"{name}_gray": gray, f"{name}_down": down, f"{name}_canny": edges,
                    f"{name}_canny_60_20": edges2, f"{name}_dt": dt})
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(os.path.join(HERE, "cv2_small.npz"), **out)


if __name__ == "__main__":
    make_se3()
    make_cv2()
    print("golden vectors written to", HERE)
