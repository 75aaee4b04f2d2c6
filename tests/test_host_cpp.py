"""The C++ host mirror (revo_b200/host/revo_host.hpp) compiles against the C ABI with plain g++, links the CUDA
library and reproduces the reference's call sequence.  CPU: build + self-test (settings/camera rules, clean
REVO_ERR_NO_DEVICE without a GPU).  GPU: track a pair from C++ and compare with the python binding."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rot_angle, synth_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_exe(tmp_path_factory):
    from revo_b200 import build

    lib = build.build()
    out = os.path.join(ROOT, "tests", "cpp", "test_host")
    src = os.path.join(ROOT, "tests", "cpp", "test_host.cpp")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        libdir = os.path.dirname(lib)
        subprocess.run(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", src, "-o", out, "-L", libdir, "-lrevo_b200",
                        f"-Wl,-rpath,{libdir}"], check=True)
    return out


def test_host_header_compiles_and_selftests(host_exe):
    r = subprocess.run([host_exe, "--selftest"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selftest ok" in r.stdout


@pytest.mark.gpu
def test_host_cpp_tracks_like_python_binding(host_exe, ctx, tmp_path):
    from revo_b200 import api, synth

    p = synth_pair(41, 320, 240)
    fx, fy, cx, cy, w, h = p["cam"]
    fix = tmp_path / "pair.bin"
    with open(fix, "wb") as f:
        f.write(np.array([w, h, 3, 0], np.int32).tobytes())
        f.write(np.array([fx, fy, cx, cy], np.float32).tobytes())
        for key in ("key", "cur"):
            f.write(np.ascontiguousarray(p[key][0]).tobytes())
            f.write(np.ascontiguousarray(p[key][1]).tobytes())
    out = tmp_path / "out.txt"
    r = subprocess.run([host_exe, str(fix), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(out).read().strip().split("\n")
    status, threw, err, n0, npx = lines[0].split()
    R = np.array([float(v) for v in lines[1].split()], np.float32).reshape(3, 3).T
    T = np.array([float(v) for v in lines[2].split()], np.float32)
    assert int(threw) == 1 and int(npx) == w * h
    st = api.ImgPyramidSettings(PYR_MIN_LVL=2, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    k = api.ImgPyramidRGBD(ctx, st, None, *p["key"])
    c = api.ImgPyramidRGBD(ctx, st, None, *p["cur"])
    k.makeKeyframe()
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    s2, R2, T2, e2 = trk.trackFrames(np.eye(3), np.zeros(3), k, c)
    assert int(status) == s2 and int(n0) == c.returnNumEdges(0)
    assert np.array_equal(R, R2) and np.array_equal(T, T2) and abs(float(err) - e2) < 1e-7
    Tgt = p["T_kf_cur"]
    assert rot_angle(R, Tgt[:3, :3]) < 2e-3 and np.linalg.norm(T - Tgt[:3, 3]) < 4e-3
