"""CPU ORACLE bindings (test infrastructure, NOT product code).

ctypes front-end for ``oracle/revo_oracle.c`` plus the python-side
orchestration of the reference's pyramid constructor
(``datastructures/imgpyramidrgbd.cpp:43-96,173-252``).  Two pyramid flavours:

* ``backend="cv2"``  -- the four OpenCV calls of the reference go to python
  ``cv2`` 4.13 (the same OpenCV kernels the reference links, newer release;
  reference pins "OpenCV 3", ``CMakeLists.txt:46``), the hand-written loops go
  to the C restatement.  This is the parity anchor.
* ``backend="c"``    -- everything through the dependency-free C restatement;
  ``tests/test_oracle.py`` proves it bit-identical to the cv2 flavour.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(force: bool = False) -> None:
    """Compile the C oracle (``make -C oracle``)."""
    need = force or not all(
        os.path.exists(os.path.join(_BUILD, f"librevo_oracle_{p}.so")) for p in ("f32", "f64")
    )
    if not need:
        src = os.path.getmtime(os.path.join(_HERE, "revo_oracle.c"))
        need = any(os.path.getmtime(os.path.join(_BUILD, f"librevo_oracle_{p}.so")) < src for p in ("f32", "f64"))
    if need:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)


class OptCfg(C.Structure):
    """``OptimizerSettings`` -- system/optimizer.h:46-111."""

    _fields_ = [
        ("lambda_success_fac", C.c_float),
        ("lambda_fail_fac", C.c_float),
        ("lambda_initial", C.c_float * 6),
        ("step_size_min", C.c_float * 6),
        ("convergence_eps", C.c_float * 6),
        ("max_its_per_lvl", C.c_int * 6),
        ("edge_distance_lvl", C.c_float * 6),
        ("huber_edge", C.c_float),
        ("use_edge_filter", C.c_int),
    ]


class Cam(C.Structure):
    """``Camera`` -- datastructures/camerapyr.h:90-111."""

    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("w", C.c_int), ("h", C.c_int)]


class TraceEntry(C.Structure):
    _fields_ = [("error", C.c_float), ("lam", C.c_float), ("accepted", C.c_int), ("good", C.c_int), ("bad", C.c_int)]


def _mk_resinfo(real):
    class ResInfo(C.Structure):
        _fields_ = [("good", C.c_int), ("bad", C.c_int), ("sum_w", real), ("sum_unw", real)]

    return ResInfo


def _mk_level():
    class Level(C.Structure):
        _fields_ = [
            ("cur_pts4", C.c_void_p),
            ("cur_n", C.c_int),
            ("ref_opt4", C.c_void_p),
            ("ref_dt", C.c_void_p),
            ("cam", Cam),
        ]

    return Level


Level = _mk_level()


def level_cam(fx, fy, cx, cy, w, h, lvl) -> Cam:
    """``Camera(fx,fy,cx,cy,w,h,scale)`` -- camerapyr.h:98-103 with
    ``scale = 1.0f/pow(2,lvl)`` (:141); level 0 is unscaled (:138)."""
    if lvl == 0:
        return Cam(np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy), int(w), int(h))
    s = np.float32(1.0 / (2.0 ** lvl))
    f32 = np.float32
    return Cam(f32(fx) * s, f32(fy) * s, f32(cx) * s, f32(cy) * s, int(f32(w) * s), int(f32(h) * s))


class Oracle:
    """One precision flavour of the C oracle ("f32" = reference-as-is, "f64" = truth)."""

    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        build()
        self.precision = precision
        self.real = C.c_float if precision == "f32" else C.c_double
        self.np_real = np.float32 if precision == "f32" else np.float64
        self.ResInfo = _mk_resinfo(self.real)
        self.lib = C.CDLL(os.path.join(_BUILD, f"librevo_oracle_{precision}.so"))
        L = self.lib
        assert L.orc_sizeof_real() == C.sizeof(self.real)
        L.orc_track_level.restype = self.real
        L.orc_eval_cost_function.restype = self.real
        L.orc_dist_histogram.restype = C.c_float
        L.orc_edges3d.restype = C.c_int
        L.orc_track_frames.restype = C.c_int

    def set_num_threads(self, n: int) -> int:
        """Threads of the OpenMP batch harness (``track_frames_batch``); returns the count in effect."""
        self.lib.orc_set_num_threads.restype = C.c_int
        return int(self.lib.orc_set_num_threads(C.c_int(int(n))))

    # -- helpers ---------------------------------------------------------
    def _r(self, a):
        return np.ascontiguousarray(np.asarray(a, dtype=self.np_real))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def default_cfg(self) -> OptCfg:
        c = OptCfg()
        self.lib.orc_opt_cfg_default(C.byref(c))
        return c

    # -- SE3 -------------------------------------------------------------
    def se3_exp(self, xi):
        xi = self._r(xi)
        q = np.zeros(4, self.np_real)
        t = np.zeros(3, self.np_real)
        self.lib.orc_se3_exp(self._p(xi), self._p(q), self._p(t))
        return q, t

    def se3_from_Rt(self, R, t):
        """R is a 3x3 numpy matrix (row-major view); converted to Eigen's column-major."""
        Rc = self._r(np.asarray(R).T.reshape(-1))
        t = self._r(t)
        q = np.zeros(4, self.np_real)
        to = np.zeros(3, self.np_real)
        rc = self.lib.orc_se3_from_Rt(self._p(Rc), self._p(t), self._p(q), self._p(to))
        return rc, q, to

    def se3_mul(self, qa, ta, qb, tb):
        qa, ta, qb, tb = map(self._r, (qa, ta, qb, tb))
        q = np.zeros(4, self.np_real)
        t = np.zeros(3, self.np_real)
        self.lib.orc_se3_mul(self._p(qa), self._p(ta), self._p(qb), self._p(tb), self._p(q), self._p(t))
        return q, t

    def quat_to_R(self, q):
        q = self._r(q)
        R = np.zeros(9, self.np_real)
        self.lib.orc_quat_to_R(self._p(q), self._p(R))
        return R.reshape(3, 3).T.copy()

    def ldlt_solve6(self, A, b):
        A = self._r(np.asarray(A).T.reshape(-1))  # column-major
        b = self._r(b)
        x = np.zeros(6, self.np_real)
        self.lib.orc_ldlt_solve6(self._p(A), self._p(b), self._p(x))
        return x

    # -- optimizer ---------------------------------------------------------
    def eval_record(self, pts4, opt4, cam: Cam, R, T, cfg: OptCfg, lvl: int) -> np.ndarray:
        pts4 = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
        opt4 = np.ascontiguousarray(opt4, np.float32)
        Rc = self._r(np.asarray(R).T.reshape(-1))
        T = self._r(T)
        rec = np.zeros(32, np.float64)
        self.lib.orc_eval_record(self._p(pts4), C.c_int(len(pts4)), self._p(opt4), C.byref(cam), self._p(Rc),
                                 self._p(T), C.byref(cfg), C.c_int(lvl), self._p(rec))
        return rec

    def track_level(self, pts4, opt4, cam: Cam, R, T, cfg: OptCfg, lvl: int, max_tries: int = 0, trace_cap: int = 4096):
        """``Optimizer::trackFrames``. Returns dict(R, T, error, good, bad, sum_w, sum_unw, n_evals, trace, rc)."""
        pts4 = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
        opt4 = np.ascontiguousarray(opt4, np.float32)
        Rc = self._r(np.asarray(R).T.reshape(-1)).copy()
        Tc = self._r(T).copy()
        ri = self.ResInfo()
        trace = (TraceEntry * trace_cap)()
        n_evals = C.c_int(0)
        rc = C.c_int(0)
        err = self.lib.orc_track_level(self._p(pts4), C.c_int(len(pts4)), self._p(opt4), C.byref(cam), self._p(Rc),
                                       self._p(Tc), C.byref(cfg), C.c_int(lvl), C.byref(ri), trace, C.c_int(trace_cap),
                                       C.byref(n_evals), C.c_int(max_tries), C.byref(rc))
        ntr = max(0, n_evals.value - 1)
        tr = [(trace[i].error, trace[i].lam, trace[i].accepted, trace[i].good, trace[i].bad) for i in range(min(ntr, trace_cap))]
        return dict(R=Rc.reshape(3, 3).T.copy(), T=Tc.copy(), error=float(err), good=ri.good, bad=ri.bad,
                    sum_w=float(ri.sum_w), sum_unw=float(ri.sum_unw), n_evals=n_evals.value, trace=tr, rc=rc.value)

    def eval_cost_function(self, pts4, dt, cam: Cam, R, T, cfg: OptCfg, lvl: int) -> float:
        pts4 = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
        dt = np.ascontiguousarray(dt, np.float32)
        Rc = self._r(np.asarray(R).T.reshape(-1))
        T = self._r(T)
        return float(self.lib.orc_eval_cost_function(self._p(pts4), C.c_int(len(pts4)), self._p(dt), C.byref(cam),
                                                     self._p(Rc), self._p(T), C.byref(cfg), C.c_int(lvl)))

    def _levels_array(self, ref: "Pyramid", cur: "Pyramid", keep: list):
        arr = (Level * 6)()
        for l in range(min(6, cur.n_levels)):
            p = np.ascontiguousarray(cur.edges3d[l], np.float32)
            o = np.ascontiguousarray(ref.opt[l], np.float32)
            d = np.ascontiguousarray(ref.dt[l], np.float32)
            keep += [p, o, d]
            arr[l].cur_pts4 = p.ctypes.data
            arr[l].cur_n = len(p)
            arr[l].ref_opt4 = o.ctypes.data
            arr[l].ref_dt = d.ctypes.data
            arr[l].cam = cur.cams[l]
        return arr

    def track_frames(self, ref: "Pyramid", cur: "Pyramid", R, T, cfg: OptCfg, min_lvl: int, max_lvl: int = 0,
                     check_init: bool = True):
        """``TrackerNew::trackFrames`` (system/tracker.cpp:294-353)."""
        keep: list = []
        arr = self._levels_array(ref, cur, keep)
        Rc = self._r(np.asarray(R).T.reshape(-1)).copy()
        Tc = self._r(T).copy()
        err = self.real(0)
        ri = self.ResInfo()
        evals = (C.c_int * 6)()
        rc = C.c_int(0)
        status = self.lib.orc_track_frames(arr, C.c_int(min_lvl), C.c_int(max_lvl), C.c_int(int(check_init)),
                                           C.byref(cfg), self._p(Rc), self._p(Tc), C.byref(err), C.byref(ri), evals,
                                           C.byref(rc))
        return dict(R=Rc.reshape(3, 3).T.copy(), T=Tc.copy(), error=float(err.value), status=status, good=ri.good,
                    bad=ri.bad, evals=list(evals), rc=rc.value)

    def track_frames_traced(self, ref: "Pyramid", cur: "Pyramid", R, T, cfg: OptCfg, min_lvl: int, max_lvl: int = 0,
                            check_init: bool = True, trace_cap: int = 1024):
        """``track_frames`` plus the accept / reject sequence of every level: ``accepts[lvl]`` is a string of 'A' / 'r', one
        character per LM try (coarse-to-fine order of the levels is the caller's to impose)."""
        keep: list = []
        arr = self._levels_array(ref, cur, keep)
        Rc = self._r(np.asarray(R).T.reshape(-1)).copy()
        Tc = self._r(T).copy()
        err = self.real(0)
        ri = self.ResInfo()
        evals = (C.c_int * 6)()
        per = (C.c_int * 6)()
        rc = C.c_int(0)
        trace = (TraceEntry * trace_cap)()
        self.lib.orc_track_frames_traced.restype = C.c_int
        status = self.lib.orc_track_frames_traced(arr, C.c_int(min_lvl), C.c_int(max_lvl), C.c_int(int(check_init)), C.byref(cfg),
                                                  self._p(Rc), self._p(Tc), C.byref(err), C.byref(ri), evals, C.byref(rc), trace,
                                                  C.c_int(trace_cap), per)
        accepts, k = {}, 0
        for lvl in range(min_lvl, max_lvl - 1, -1):
            accepts[lvl] = "".join("A" if trace[k + i].accepted else "r" for i in range(per[lvl]))
            k += per[lvl]
        return dict(R=Rc.reshape(3, 3).T.copy(), T=Tc.copy(), error=float(err.value), status=status, good=ri.good,
                    bad=ri.bad, evals=list(evals), rc=rc.value, accepts=accepts)

    def track_frames_batch(self, refs, curs, Rs, Ts, cfg: OptCfg, min_lvl: int, max_lvl: int = 0, check_init: bool = True):
        """OpenMP batch over independent pairs (CPU-baseline harness)."""
        n = len(refs)
        keep: list = []
        arr = (Level * (6 * n))()
        for p in range(n):
            a = self._levels_array(refs[p], curs[p], keep)
            for l in range(6):
                arr[6 * p + l] = a[l]
        Rc = np.ascontiguousarray(np.stack([np.asarray(R).T.reshape(-1) for R in Rs]).astype(self.np_real))
        Tc = np.ascontiguousarray(np.stack(Ts).astype(self.np_real))
        errs = np.zeros(n, self.np_real)
        status = np.zeros(n, np.int32)
        evals = np.zeros((n, 6), np.int32)
        self.lib.orc_track_frames_batch(arr, C.c_int(n), C.c_int(min_lvl), C.c_int(max_lvl), C.c_int(int(check_init)),
                                        C.byref(cfg), self._p(Rc), self._p(Tc), self._p(errs), self._p(status),
                                        self._p(evals))
        Rout = Rc.reshape(n, 3, 3).transpose(0, 2, 1).copy()
        return dict(R=Rout, T=Tc, error=errs, status=status, evals=evals)

    # -- pyramid pieces ----------------------------------------------------
    def subsample_depth(self, depth):
        depth = np.ascontiguousarray(depth, np.float32)
        h, w = depth.shape
        out = np.zeros((h // 2, w // 2), np.float32)
        self.lib.orc_subsample_depth_holes(self._p(depth), C.c_int(w), C.c_int(h), self._p(out))
        return out

    def dist_histogram(self, edges, P):
        edges = np.ascontiguousarray(edges, np.uint8)
        h, w = edges.shape
        hist = np.zeros((h // P, w // P), np.uint8)
        frac = self.lib.orc_dist_histogram(self._p(edges), C.c_int(w), C.c_int(h), C.c_int(P), self._p(hist))
        return hist, float(frac)

    def fill_in_edges(self, top, hist, P, P_low, edges_mod):
        top = np.ascontiguousarray(top, np.uint8)
        hist = np.ascontiguousarray(hist, np.uint8)
        assert edges_mod.dtype == np.uint8 and edges_mod.flags.c_contiguous
        self.lib.orc_fill_in_edges(self._p(top), C.c_int(top.shape[1]), C.c_int(top.shape[0]), self._p(hist),
                                   C.c_int(hist.shape[1]), C.c_int(hist.shape[0]), C.c_int(P), C.c_int(P_low),
                                   self._p(edges_mod), C.c_int(edges_mod.shape[1]), C.c_int(edges_mod.shape[0]))

    def edges3d(self, edges, depth, cam: Cam, dmin, dmax):
        edges = np.ascontiguousarray(edges, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        out = np.zeros((cam.w * cam.h, 4), np.float32)
        n = self.lib.orc_edges3d(self._p(edges), self._p(depth), C.byref(cam), C.c_float(dmin), C.c_float(dmax), self._p(out))
        return out[:n].copy()

    def build_opt_structure(self, dt):
        dt = np.ascontiguousarray(dt, np.float32)
        h, w = dt.shape
        out = np.zeros((h, w, 4), np.float32)
        self.lib.orc_build_opt_structure(self._p(dt), C.c_int(w), C.c_int(h), self._p(out))
        return out

    def gray_bgr(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, ch = bgr.shape
        out = np.zeros((h, w), np.uint8)
        self.lib.orc_gray_bgr(self._p(bgr), C.c_int(w), C.c_int(h), C.c_size_t(w * ch), C.c_int(ch), self._p(out))
        return out

    def pyrdown(self, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
        self.lib.orc_pyrdown_u8(self._p(gray), C.c_int(w), C.c_int(h), self._p(out))
        return out

    def canny(self, gray, t1, t2):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        out = np.zeros((h, w), np.uint8)
        self.lib.orc_canny(self._p(gray), C.c_int(w), C.c_int(h), C.c_double(t1), C.c_double(t2), self._p(out))
        return out

    def edt(self, edges):
        edges = np.ascontiguousarray(edges, np.uint8)
        h, w = edges.shape
        out = np.zeros((h, w), np.float32)
        self.lib.orc_edt_l2(self._p(edges), C.c_int(w), C.c_int(h), self._p(out))
        return out


# ---------------------------------------------------------------------------
# Pyramid orchestration (imgpyramidrgbd.cpp:43-96, 173-229, 231-252)
# ---------------------------------------------------------------------------
@dataclass
class PyrCfg:
    """``ImgPyramidSettings`` -- datastructures/camerapyr.h:27-89 (hot-path fields)."""

    n_levels: int = 3
    canny1: int = 150
    canny2: int = 100
    depth_min: float = 0.1
    depth_max: float = 5.2
    use_edge_hist: bool = True
    n_percentage: float = 0.3
    # distPatchSizes (imgpyramidrgbd.cpp:50) has 3 entries and the reference
    # indexes it out of bounds for level >= 3.  Patch sizes are a
    # function of the level-0 size so that the grid is always 32x24-like:
    # P_l = (w0/32) >> l ; levels whose P_l < 1... are given P=max(1,..) and
    # never filled in (fill-in only defined for the reference's 3 levels).
    patch0: int = 20


@dataclass
class Pyramid:
    n_levels: int
    cams: list
    gray: list = field(default_factory=list)
    depth: list = field(default_factory=list)
    edges: list = field(default_factory=list)       # after fill-in (edgesPyr)
    edges_orig: list = field(default_factory=list)  # Canny output (edgesOrigPyr)
    hist: list = field(default_factory=list)
    filled: list = field(default_factory=list)
    edges3d: list = field(default_factory=list)     # column-major order, (N,4) f32
    dt: list = field(default_factory=list)          # keyframe only
    opt: list = field(default_factory=list)         # keyframe only (h,w,4) f32


def patch_size(cfg: PyrCfg, lvl: int) -> int:
    """distPatchSizes = {20,10,5} (imgpyramidrgbd.cpp:50), extended as patch0 >> lvl (min 1)."""
    return max(1, cfg.patch0 >> lvl)


def build_pyramid(orc: Oracle, cfg: PyrCfg, cam0, bgr, depth, backend: str = "cv2") -> Pyramid:
    """``ImgPyramidRGBD::ImgPyramidRGBD`` (imgpyramidrgbd.cpp:43-96).  cam0 = (fx,fy,cx,cy,w,h)."""
    fx, fy, cx, cy, w, h = cam0
    cams = [level_cam(fx, fy, cx, cy, w, h, l) for l in range(cfg.n_levels)]
    pyr = Pyramid(cfg.n_levels, cams)
    if backend == "cv2":
        import cv2

        gray = cv2.cvtColor(np.ascontiguousarray(bgr), cv2.COLOR_BGR2GRAY if bgr.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)  # :53
    else:
        gray = orc.gray_bgr(bgr)
    d = np.ascontiguousarray(depth, np.float32).copy()  # :54
    for lvl in range(cfg.n_levels):
        if lvl > 0:
            if backend == "cv2":
                import cv2

                gray = cv2.pyrDown(gray)  # :82
            else:
                gray = orc.pyrdown(gray)
            d = orc.subsample_depth(d)  # :84
        _add_level_edge(orc, cfg, pyr, gray, d, cams[lvl], lvl, backend)
    return pyr


def _add_level_edge(orc, cfg, pyr, gray, depth, cam, lvl, backend):
    """``addLevelEdge`` (imgpyramidrgbd.cpp:173-229)."""
    pyr.gray.append(gray)
    pyr.depth.append(depth)
    if backend == "cv2":
        import cv2

        edges = cv2.Canny(gray, cfg.canny1, cfg.canny2, apertureSize=3, L2gradient=True)  # :184
    else:
        edges = orc.canny(gray, cfg.canny1, cfg.canny2)
    pyr.edges_orig.append(edges.copy())
    edges = np.ascontiguousarray(edges)
    P = patch_size(cfg, lvl)
    hist, frac = orc.dist_histogram(edges, P)  # :187
    pyr.hist.append(hist)
    filled = False
    if cfg.use_edge_hist and lvl >= 1 and lvl <= 2:  # :188 (levels >= 3: reference UB, SURVEY D5 -> never filled)
        if np.float32(frac) < np.float32(cfg.n_percentage):  # :192
            orc.fill_in_edges(pyr.edges[lvl - 1], hist, P, patch_size(cfg, lvl - 1), edges)
            filled = True
    pyr.filled.append(filled)
    pyr.edges.append(edges)
    pyr.edges3d.append(orc.edges3d(edges, depth, cam, cfg.depth_min, cfg.depth_max))  # :199-226


def cv2_distance_transform(edges: np.ndarray) -> np.ndarray:
    """``cv::distanceTransform(255-edges, CV_DIST_L2, CV_DIST_MASK_PRECISE)`` through OpenCV's OWN
    implementation (trueDistTrans).  cv2 4.13 routes images with fewer than 2^14 pixels (or any image when
    it runs single-threaded) to Intel IPP instead (distransform.cpp, IPP_DISABLE_PERF_TRUE_DIST_MT), whose
    square root is 1 ulp off the correctly rounded value on some inputs; every VGA-or-larger level goes
    through trueDistTrans = exactly sqrtf(d^2).  The parity target is that exact definition, so the IPP
    branch is switched off around this one call."""
    import cv2

    had = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        return cv2.distanceTransform(255 - edges, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    finally:
        cv2.ipp.setUseIPP(had)


def make_keyframe(orc: Oracle, pyr: Pyramid, backend: str = "cv2") -> None:
    """``ImgPyramidRGBD::makeKeyframe`` (imgpyramidrgbd.cpp:231-252)."""
    pyr.dt, pyr.opt = [], []
    for lvl in range(pyr.n_levels):
        if backend == "cv2":
            import cv2

            dt = cv2_distance_transform(pyr.edges[lvl])  # :241
        else:
            dt = orc.edt(pyr.edges[lvl])
        pyr.dt.append(dt)
        pyr.opt.append(orc.build_opt_structure(dt))  # :245


def assess_tracking_quality(past_pts, past_world_poses, estimated_pose, cam, depth, edges_orig, depth_min=0.1, depth_max=5.2,
                            n_frames_voting=3):
    """CPU restatement (numpy, float32 in the reference's operation order) of TrackerNew::assessTrackingQuality,
    system/tracker.cpp:118-201.  past_pts: list of (N_i, 4) float32 3-D edge lists (return3DEdges(histogramLevel)),
    past_world_poses: list of 4x4, cam = (fx, fy, cx, cy, w, h) of the histogram level, depth / edges_orig of the current frame
    at that level.  Returns dict(histogram, overlaps, overlap_measure, status, out_of_bounds)."""
    f32 = np.float32
    fx, fy, cx, cy, w, h = cam
    fx, fy, cx, cy = f32(fx), f32(fy), f32(cx), f32(cy)
    w, h = int(w), int(h)
    nf = min(len(past_pts), int(n_frames_voting), 3)
    M = np.zeros((h, w), np.int32)
    oob = 0
    est_inv = np.linalg.inv(np.asarray(estimated_pose, np.float32).astype(np.float64))
    for f in range(nf):
        tr = (est_inv @ np.asarray(past_world_poses[f], np.float32).astype(np.float64)).astype(f32)      # tracker.cpp:147
        R, T = tr[:3, :3], tr[:3, 3]
        p = np.asarray(past_pts[f], f32)
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        X = ((R[0, 0] * x + R[0, 1] * y) + R[0, 2] * z) + T[0]
        Y = ((R[1, 0] * x + R[1, 1] * y) + R[1, 2] * z) + T[1]
        Z = ((R[2, 0] * x + R[2, 1] * y) + R[2, 2] * z) + T[2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = (fx * X) / Z + cx                                                                          # :157-158
            v = (fy * Y) / Z + cy
        inb = (u >= 0) & (u < f32(w)) & (v >= 0) & (v < f32(h))
        oob += int((~inb).sum())
        Mi = np.zeros((h, w), np.int32)
        Mi[np.floor(v[inb]).astype(np.int64), np.floor(u[inb]).astype(np.int64)] = 1                      # one mark per frame and pixel
        M += Mi
    d = np.asarray(depth, f32)
    ok = np.isfinite(d) & (d > f32(depth_min)) & (d < f32(depth_max))                                     # isPointOkDepth
    e = np.asarray(edges_orig) > 0
    histogram = [int((ok & (M == k)).sum()) for k in range(4)]
    overlaps = [int((ok & e & (M == k)).sum()) for k in range(4)]
    weights = [0.0, 1.0, 1.25, 1.5]
    measure = float(sum(f32(overlaps[k]) * f32(weights[k]) for k in range(1, nf + 1)))
    status = 0 if (measure >= overlaps[0] or nf + 1 < 4) else 2                                            # OK / NEW_KF
    return dict(histogram=histogram, overlaps=overlaps, overlap_measure=measure, status=status, out_of_bounds=oob, n_frames=nf)


def generate_colored_pcl(bgr, depth, edges, cam, depth_min=0.1, depth_max=5.2, dense=False):
    """``ImgPyramidRGBD::generateColoredPcl`` loop (datastructures/imgpyramidrgbd.cpp:300-323) restated literally (pure
    Python loops: small cases only).  cam = (fx, fy, cx, cy, w, h); bgr is the colour image already reduced to the level
    (:287-296).  Returns (8, N) float32 columns (X, Y, Z, 1, r, g, b, 1) in the reference's scan order (x outer, y inner)."""
    fx, fy, cx, cy = (np.float32(v) for v in cam[:4])
    cols = []
    h, w = depth.shape
    for xx in range(w):
        for yy in range(h):
            Z = np.float32(depth[yy, xx])
            if np.isfinite(Z) and Z > np.float32(depth_min) and Z < np.float32(depth_max):     # isPointOkDepth, imgpyramidrgbd.h:170-173
                if dense or edges[yy, xx] > 0:
                    b, g, r = (np.float32(v) for v in bgr[yy, xx][:3])
                    X = Z * (np.float32(xx) - cx) / fx
                    Y = Z * (np.float32(yy) - cy) / fy
                    cols.append([X, Y, Z, 1.0, r / np.float32(255.0), g / np.float32(255.0), b / np.float32(255.0), 1.0])
    return np.asarray(cols, np.float32).reshape(-1, 8).T.copy()
