"""ctypes front-end of ``oracle/_ref/librevo_ref.so``: the REFERENCE'S OWN hot-path classes (``ImgPyramidRGBD``,
``Optimizer``, ``TrackerNew``, ``LGS6``) compiled from ``/root/reference`` against the API shims of ``oracle/shim/``
(see ``oracle/ref_harness.cpp``).  TEST INFRASTRUCTURE: it pins the restatement in ``oracle/revo_oracle.c`` /
``oracle/oracle.py`` -- the checker of the CUDA path -- and generates the golden vectors of ``tests/golden/ref_*.npz``.
Only ``tests/`` may import it."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

from .oracle import OptCfg

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "librevo_ref.so")
REFERENCE = os.environ.get("REVO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(LIB)


def build(force: bool = False) -> bool:
    """Compile oracle/_ref from the reference's sources where they lie (only possible where /root/reference exists)."""
    if not os.path.isdir(os.path.join(REFERENCE, "system")):
        return available()
    subprocess.run(["make", "-C", _HERE, "-s", "ref", f"REF={REFERENCE}"] + (["-B"] if force else []), check=True)
    return available()


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.ref_pyr_create.restype = vp
        L.ref_pyr_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_float, C.c_int, C.c_float, vp, C.c_int, vp, C.c_double]
        L.ref_pyr_make_keyframe.argtypes = [vp]
        L.ref_pyr_destroy.argtypes = [vp]
        L.ref_pyr_level_size.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
        L.ref_pyr_num_edges.argtypes = [vp, C.c_int]
        L.ref_pyr_get.restype = C.c_long
        L.ref_pyr_get.argtypes = [vp, C.c_int, C.c_int, vp, C.c_long]
        L.ref_pyr_colored_pcl.restype = C.c_long
        L.ref_pyr_colored_pcl.argtypes = [vp, C.c_int, C.c_int, vp, C.c_long]
        L.ref_opt_track_level.restype = C.c_float
        L.ref_opt_track_level.argtypes = [vp, vp, C.POINTER(OptCfg), C.c_int, vp, vp] + [vp] * 4
        L.ref_opt_eval.restype = C.c_float
        L.ref_opt_eval.argtypes = [vp, vp, C.POINTER(OptCfg), C.c_int, vp, vp] + [vp] * 7
        L.ref_tracker_create.restype = vp
        L.ref_tracker_create.argtypes = [vp, C.POINTER(OptCfg), C.c_int, C.c_int, C.c_int]
        L.ref_tracker_destroy.argtypes = [vp]
        L.ref_tracker_track.argtypes = [vp, vp, vp, vp, vp, vp]
        L.ref_tracker_eval_cost.restype = C.c_float
        L.ref_tracker_eval_cost.argtypes = [vp, vp, vp, vp, vp, C.c_int]
        L.ref_tracker_add_old.argtypes = [vp, vp, vp, C.c_double]
        L.ref_tracker_clear_past.argtypes = [vp]
        L.ref_tracker_num_past.argtypes = [vp]
        L.ref_tracker_assess.argtypes = [vp, vp, vp]
        L.ref_lgs6.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int]
        L.ref_interp43.argtypes = [vp, C.c_int, C.c_float, C.c_float, vp]
        L.ref_last_log.restype = C.c_long
        L.ref_last_log.argtypes = [vp, C.c_long]
        _lib = L
    return _lib


def last_log() -> str:
    n = lib().ref_last_log(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib().ref_last_log(buf, n + 1)
    return buf.value.decode(errors="replace")


def lm_trace_from_log(log: str):
    """The LM tries the reference logged (optimizer.cpp:270): list of (good, bad, error) in evaluation order."""
    out = []
    for m in re.finditer(r"goodPts: (\d+) bad: (\d+) tot: \d+ error = ([-+0-9.eEnaif]+)=", log):
        out.append((int(m.group(1)), int(m.group(2)), float(m.group(3))))
    return out


_GET = {"gray": (0, np.uint8), "depth": (1, np.float32), "edges": (2, np.uint8), "edges_orig": (3, np.uint8), "hist": (4, np.uint8),
        "edges3d": (5, np.float32), "dt": (6, np.float32), "opt": (7, np.float32)}


class RefPyramid:
    """``ImgPyramidRGBD(settings, cameraPyr, rgb, depth, ts)`` of the reference (imgpyramidrgbd.cpp:43-96)."""

    def __init__(self, cam, n_levels, bgr, depth, canny=(150, 100), dmin=0.1, dmax=5.2, use_hist=True, n_percentage=0.3, ts=0.0):
        fx, fy, cx, cy, w, h = cam
        bgr = np.ascontiguousarray(bgr, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        self.n_levels = n_levels
        self.h = lib().ref_pyr_create(int(w), int(h), fx, fy, cx, cy, n_levels, canny[0], canny[1], dmin, dmax, int(use_hist), n_percentage,
                                      bgr.ctypes.data, bgr.shape[2], depth.ctypes.data, ts)

    def make_keyframe(self):
        lib().ref_pyr_make_keyframe(self.h)

    def level_size(self, lvl):
        w, h = C.c_int(), C.c_int()
        cam4 = np.zeros(4, np.float32)
        lib().ref_pyr_level_size(self.h, lvl, C.byref(w), C.byref(h), cam4.ctypes.data)
        return w.value, h.value, cam4

    def get(self, what: str, lvl: int) -> np.ndarray:
        which, dt = _GET[what]
        n = lib().ref_pyr_get(self.h, lvl, which, None, 0)
        assert n >= 0, what
        buf = np.zeros(n // np.dtype(dt).itemsize, dt)
        lib().ref_pyr_get(self.h, lvl, which, buf.ctypes.data, n)
        w, h, _ = self.level_size(lvl)
        if what == "edges3d":
            return buf.reshape(-1, 4)
        if what == "opt":
            return buf.reshape(h, w, 4)
        if what == "hist":
            return buf          # (h/P) x (w/P), caller reshapes
        return buf.reshape(h, w)

    def colored_pcl(self, lvl: int, dense: bool) -> np.ndarray:
        n = lib().ref_pyr_colored_pcl(self.h, lvl, int(dense), None, 0)
        buf = np.zeros((n, 8), np.float32)
        lib().ref_pyr_colored_pcl(self.h, lvl, int(dense), buf.ctypes.data, buf.size)
        return buf.T.copy()       # (8, N)

    def close(self):
        if self.h:
            lib().ref_pyr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pose_args(R, T):
    Rc = np.ascontiguousarray(np.asarray(R, np.float32).reshape(3, 3).T.reshape(-1)).copy()      # Eigen column-major
    Tc = np.ascontiguousarray(np.asarray(T, np.float32).reshape(3)).copy()
    return Rc, Tc


def opt_track_level(ref: RefPyramid, cur: RefPyramid, cfg: OptCfg, lvl: int, R, T):
    """``Optimizer::trackFrames`` (optimizer.cpp:235-311).  The LM trace is parsed from the reference's own log."""
    Rc, Tc = _pose_args(R, T)
    good, bad = C.c_int(), C.c_int()
    sw, su = C.c_float(), C.c_float()
    err = lib().ref_opt_track_level(ref.h, cur.h, C.byref(cfg), lvl, Rc.ctypes.data, Tc.ctypes.data, C.addressof(good), C.addressof(bad),
                                    C.addressof(sw), C.addressof(su))
    trace = lm_trace_from_log(last_log())
    return dict(R=Rc.reshape(3, 3).T.copy(), T=Tc, error=float(err), good=good.value, bad=bad.value, sum_w=sw.value, sum_unw=su.value,
                trace=trace, n_evals=1 + len(trace))


def opt_eval(ref: RefPyramid, cur: RefPyramid, cfg: OptCfg, lvl: int, R, T):
    """calcErrorAndBuffers + calculateWarpUpdate + LGS6::finish at one pose."""
    Rc, Tc = _pose_args(R, T)
    good, bad = C.c_int(), C.c_int()
    sw, su, lse = C.c_float(), C.c_float(), C.c_float()
    A, b = np.zeros(36, np.float32), np.zeros(6, np.float32)
    err = lib().ref_opt_eval(ref.h, cur.h, C.byref(cfg), lvl, Rc.ctypes.data, Tc.ctypes.data, C.addressof(good), C.addressof(bad),
                             C.addressof(sw), C.addressof(su), A.ctypes.data, b.ctypes.data, C.addressof(lse))
    return dict(error=float(err), good=good.value, bad=bad.value, sum_w=sw.value, sum_unw=su.value, A=A.reshape(6, 6).T.copy(), b=b,
                ls_error=lse.value)


class RefTracker:
    """``TrackerNew`` of the reference (tracker.cpp)."""

    def __init__(self, any_pyr: RefPyramid, cfg: OptCfg, check_init=True, check_tracking=True, n_frames_voting=3):
        self.h = lib().ref_tracker_create(any_pyr.h, C.byref(cfg), int(check_init), int(check_tracking), n_frames_voting)

    def track_frames(self, ref: RefPyramid, cur: RefPyramid, R, T):
        Rc, Tc = _pose_args(R, T)
        err = C.c_float()
        st = lib().ref_tracker_track(self.h, ref.h, cur.h, Rc.ctypes.data, Tc.ctypes.data, C.addressof(err))
        return dict(status=st, R=Rc.reshape(3, 3).T.copy(), T=Tc, error=err.value, trace=lm_trace_from_log(last_log()))

    def eval_cost(self, ref, cur, R, T, min_lvl):
        Rc, Tc = _pose_args(R, T)
        return float(lib().ref_tracker_eval_cost(self.h, ref.h, cur.h, Rc.ctypes.data, Tc.ctypes.data, min_lvl))

    def add_old(self, pyr: RefPyramid, world_pose, ts=0.0):
        P = np.ascontiguousarray(np.asarray(world_pose, np.float32).reshape(4, 4).T.reshape(-1))
        lib().ref_tracker_add_old(self.h, pyr.h, P.ctypes.data, ts)

    def clear_past(self):
        lib().ref_tracker_clear_past(self.h)

    def num_past(self):
        return lib().ref_tracker_num_past(self.h)

    def assess(self, cur: RefPyramid, estimated_pose) -> dict:
        """``assessTrackingQuality`` (tracker.cpp:118-201); the counts are read from the reference's own log lines."""
        P = np.ascontiguousarray(np.asarray(estimated_pose, np.float32).reshape(4, 4).T.reshape(-1))
        st = int(lib().ref_tracker_assess(self.h, cur.h, P.ctypes.data))
        log = last_log()
        hist = [(int(a), int(b), int(c)) for a, b, c in re.findall(r"histLvl: (\d+) total (\d+) overlap: (\d+)", log)]
        oob = re.search(r"outOfBounds: (\d+)", log)
        meas = re.search(r"overlapMeasure: ([-+0-9.eE]+)", log)
        return dict(status=st, histogram=[t for _, t, _ in hist], overlaps=[o for _, _, o in hist],
                    out_of_bounds=int(oob.group(1)) if oob else None, overlap_measure=float(meas.group(1)) if meas else None)

    def close(self):
        if self.h:
            lib().ref_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lgs6(J, res, w, finish=True):
    J = np.ascontiguousarray(J, np.float32).reshape(-1, 6)
    res = np.ascontiguousarray(res, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    A, b, e = np.zeros(36, np.float32), np.zeros(6, np.float32), C.c_float()
    lib().ref_lgs6(J.ctypes.data, res.ctypes.data, w.ctypes.data, len(J), A.ctypes.data, b.ctypes.data, C.addressof(e), int(finish))
    return A.reshape(6, 6).T.copy(), b, e.value


def interp43(opt4, x, y):
    opt4 = np.ascontiguousarray(opt4, np.float32)
    out = np.zeros(3, np.float32)
    lib().ref_interp43(opt4.ctypes.data, opt4.shape[1], x, y, out.ctypes.data)
    return out
