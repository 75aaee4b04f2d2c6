"""Host-side mirror of the reference's classes for the hot path, over the C ABI
(``include/revo_b200.h`` -> ``revo_b200/lib/librevo_b200.so``, loaded with ctypes).

Names, argument meaning and error behaviour follow the reference:

* :class:`ImgPyramidSettings`, :class:`Camera`, :class:`CameraPyr` -- datastructures/camerapyr.h
* :class:`ImgPyramidRGBD`                                        -- datastructures/imgpyramidrgbd.h:27-117
* :class:`OptimizerSettings`, :class:`Optimizer`                 -- system/optimizer.h:42-185
* :class:`TrackerSettings`, :class:`TrackerNew`                  -- system/tracker.h:31-112

There is no CPU fallback: constructing a :class:`Context` without the CUDA
library or without a GPU raises :class:`RevoError`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

MAX_LEVELS = 6

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "librevo_b200.so")
_lib = None


class RevoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"revo_b200 error {code}: {msg}")
        self.code = code


# ---- ctypes mirrors of the POD structs -------------------------------------
class revo_camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int32), ("height", C.c_int32)]


class revo_pyr_config(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("canny_threshold1", C.c_int32), ("canny_threshold2", C.c_int32),
                ("depth_min", C.c_float), ("depth_max", C.c_float), ("use_edge_hist", C.c_int32),
                ("n_percentage", C.c_float), ("patch0", C.c_int32)]


class revo_opt_config(C.Structure):
    _fields_ = [("lambda_success_fac", C.c_float), ("lambda_fail_fac", C.c_float),
                ("lambda_initial", C.c_float * MAX_LEVELS), ("step_size_min", C.c_float * MAX_LEVELS),
                ("convergence_eps", C.c_float * MAX_LEVELS), ("max_its_per_lvl", C.c_int32 * MAX_LEVELS),
                ("edge_distance_lvl", C.c_float * MAX_LEVELS), ("huber_edge", C.c_float),
                ("use_edge_filter", C.c_int32), ("max_lm_tries", C.c_int32)]


class revo_tracker_config(C.Structure):
    _fields_ = [("check_init_values", C.c_int32), ("pyr_min_lvl", C.c_int32), ("pyr_max_lvl", C.c_int32),
                ("opt", revo_opt_config)]


class revo_residual_info(C.Structure):
    _fields_ = [("good_pts_edges", C.c_int32), ("bad_pts_edges", C.c_int32),
                ("sum_error_unweighted", C.c_float), ("sum_error_weighted", C.c_float)]


class revo_track_result(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("error", C.c_float), ("status", C.c_int32),
                ("rc", C.c_int32), ("res", revo_residual_info), ("n_evals", C.c_int32 * MAX_LEVELS),
                ("n_pts", C.c_int32 * MAX_LEVELS), ("used_identity_init", C.c_int32)]


class revo_quality_result(C.Structure):
    _fields_ = [("histogram", C.c_int32 * 4), ("overlaps", C.c_int32 * 4), ("overlap_measure", C.c_float),
                ("status", C.c_int32), ("out_of_bounds", C.c_int32), ("n_frames", C.c_int32)]


QUALITY_RESULT_DTYPE = np.dtype([("histogram", np.int32, 4), ("overlaps", np.int32, 4), ("overlap_measure", np.float32), ("status", np.int32),
                                 ("out_of_bounds", np.int32), ("n_frames", np.int32)])
assert QUALITY_RESULT_DTYPE.itemsize == C.sizeof(revo_quality_result)


class revo_trace_entry(C.Structure):
    _fields_ = [("error", C.c_float), ("lam", C.c_float), ("accepted", C.c_int32), ("good", C.c_int32),
                ("bad", C.c_int32), ("level", C.c_int32)]


TRACK_RESULT_DTYPE = np.dtype([("R", np.float32, (9,)), ("t", np.float32, (3,)), ("error", np.float32),
                               ("status", np.int32), ("rc", np.int32), ("good", np.int32), ("bad", np.int32),
                               ("sum_unw", np.float32), ("sum_w", np.float32), ("n_evals", np.int32, (MAX_LEVELS,)),
                               ("n_pts", np.int32, (MAX_LEVELS,)), ("used_identity_init", np.int32)])
assert TRACK_RESULT_DTYPE.itemsize == C.sizeof(revo_track_result)

# revo_pyr_download selectors
GRAY, DEPTH, EDGES, EDGES_ORIG, HIST, EDGES3D, DT, OPTSTRUCT, EDGES3D_DEVICE_ORDER = range(9)
# TrackerNew::TrackerStatus, system/tracker.h:60-65
TRACKER_STATE_OK, TRACKER_STATE_LOST, TRACKER_STATE_NEW_KF, TRACKER_STATE_UNKNOWN = range(4)
(REVO_OK, REVO_ERR_INVALID_ARG, REVO_ERR_NO_DEVICE, REVO_ERR_CUDA, REVO_ERR_NOT_KEYFRAME, REVO_ERR_NOT_ORTHOGONAL, REVO_ERR_BAD_LEVEL,
 REVO_ERR_BUFFER_TOO_SMALL, REVO_ERR_UNSUPPORTED, REVO_ERR_COMM) = range(10)
SPLIT_HANDLE_BYTES = 128

EXPORTED_SYMBOLS = [
    "revo_pyr_config_default", "revo_opt_config_default", "revo_tracker_config_default", "revo_ctx_create",
    "revo_ctx_destroy", "revo_ctx_synchronize", "revo_strerror", "revo_last_error", "revo_ctx_stream",
    "revo_ctx_launch_count", "revo_ctx_last_timings", "revo_pyr_create", "revo_pyr_create_batch", "revo_pyr_make_keyframe",
    "revo_pyr_make_keyframe_batch", "revo_pyr_destroy", "revo_pyr_destroy_batch", "revo_pyr_is_keyframe", "revo_pyr_level_camera",
    "revo_pyr_timestamp", "revo_pyr_num_edges", "revo_pyr_download", "revo_pyr_upload_level", "revo_eval",
    "revo_track_level", "revo_track", "revo_track_batch", "revo_ctx_set_track_shape", "revo_split_export",
    "revo_split_open", "revo_track_split", "revo_ctx_set_track_engine", "revo_ctx_reserve", "revo_ctx_last_upload_ms", "revo_pyr_create_batch_u16", "revo_track_quality", "revo_quat_to_R9", "revo_R9_to_quat",
    "revo_pyr_colored_pcl", "revo_pyr_copy_points_batch", "revo_track_quality_batch", "revo_ctx_set_track_max_clusters",
]


def load_library():
    """Load the CUDA library; raises (loudly) when it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RevoError(-1, f"{_LIB_PATH} is missing: run `python -m revo_b200.build` (nvcc, sm_100a); "
                            "this package has no CPU implementation")
    lib = C.CDLL(_LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int
    lib.revo_strerror.restype = C.c_char_p
    lib.revo_last_error.restype = C.c_char_p
    lib.revo_last_error.argtypes = [vp]
    lib.revo_ctx_stream.restype = C.c_uint64
    lib.revo_ctx_stream.argtypes = [vp]
    lib.revo_ctx_launch_count.restype = C.c_uint64
    lib.revo_ctx_launch_count.argtypes = [vp]
    lib.revo_ctx_last_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.revo_pyr_timestamp.restype = C.c_double
    lib.revo_pyr_timestamp.argtypes = [vp]
    lib.revo_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.revo_ctx_destroy.argtypes = [vp]
    lib.revo_ctx_synchronize.argtypes = [vp]
    lib.revo_pyr_create.argtypes = [vp, C.POINTER(revo_pyr_config), C.POINTER(revo_camera), vp, C.c_size_t, i32, vp,
                                    C.c_size_t, C.c_double, C.POINTER(vp)]
    lib.revo_pyr_create_batch.argtypes = [vp, C.POINTER(revo_pyr_config), C.POINTER(revo_camera), i32, vp, i32, vp, vp,
                                          C.POINTER(vp)]
    lib.revo_pyr_create_batch_u16.argtypes = [vp, C.POINTER(revo_pyr_config), C.POINTER(revo_camera), i32, vp, i32, vp, C.c_float,
                                              vp, C.POINTER(vp)]
    lib.revo_pyr_make_keyframe.argtypes = [vp, vp]
    lib.revo_pyr_make_keyframe_batch.argtypes = [vp, i32, C.POINTER(vp)]
    lib.revo_pyr_destroy.argtypes = [vp, vp]
    lib.revo_pyr_destroy_batch.argtypes = [vp, i32, C.POINTER(vp)]
    lib.revo_pyr_is_keyframe.argtypes = [vp]
    lib.revo_pyr_level_camera.argtypes = [vp, i32, C.POINTER(revo_camera)]
    lib.revo_pyr_num_edges.argtypes = [vp, vp, i32, C.POINTER(i32)]
    lib.revo_pyr_download.argtypes = [vp, vp, i32, i32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.revo_pyr_upload_level.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    lib.revo_pyr_colored_pcl.argtypes = [vp, vp, i32, i32, vp, i32, vp, C.c_size_t, C.POINTER(i32)]
    lib.revo_eval.argtypes = [vp, C.POINTER(revo_opt_config), vp, vp, i32, vp, vp, vp]
    lib.revo_track_level.argtypes = [vp, C.POINTER(revo_opt_config), vp, vp, i32, vp, vp, C.POINTER(revo_residual_info),
                                     C.POINTER(C.c_float), C.POINTER(i32)]
    lib.revo_track.argtypes = [vp, C.POINTER(revo_tracker_config), vp, vp, vp, vp, C.POINTER(revo_track_result)]
    lib.revo_track_batch.argtypes = [vp, C.POINTER(revo_tracker_config), i32, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp,
                                     i32, vp]
    lib.revo_ctx_set_track_shape.argtypes = [vp, i32, i32]
    lib.revo_ctx_set_track_max_clusters.argtypes = [vp, i32]
    lib.revo_track_quality.argtypes = [vp, vp, i32, i32, C.POINTER(vp), vp, vp, i32, C.POINTER(revo_quality_result)]
    lib.revo_track_quality_batch.argtypes = [vp, i32, C.POINTER(vp), i32, vp, C.POINTER(vp), vp, vp, i32, C.POINTER(revo_quality_result)]
    lib.revo_pyr_copy_points_batch.argtypes = [vp, i32, C.POINTER(vp), i32, C.POINTER(vp)]
    lib.revo_ctx_set_track_engine.argtypes = [vp, i32, i32]
    lib.revo_ctx_reserve.argtypes = [vp, C.c_size_t]
    lib.revo_ctx_last_upload_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.revo_quat_to_R9.argtypes = [vp, vp]
    lib.revo_R9_to_quat.argtypes = [vp, vp]
    lib.revo_split_export.argtypes = [vp, i32, i32, vp]
    lib.revo_split_open.argtypes = [vp, vp]
    lib.revo_track_split.argtypes = [vp, C.POINTER(revo_tracker_config), vp, vp, vp, vp, C.POINTER(revo_track_result)]
    _lib = lib
    return lib


# ---------------------------------------------------------------------------
# settings (same names and defaults as the reference)
# ---------------------------------------------------------------------------
@dataclass
class ImgPyramidSettings:
    """datastructures/camerapyr.h:27-89 (fields of the hot path; YAML parsing is out of scope)."""
    cannyThreshold1: int = 150
    cannyThreshold2: int = 100
    DEPTH_MIN: float = 0.1
    DEPTH_MAX: float = 5.2
    PYR_MIN_LVL: int = 2
    PYR_MAX_LVL: int = 0
    width: int = 640
    height: int = 480
    fx: float = 560.0
    fy: float = 560.0
    cx: float = 320.0
    cy: float = 240.0
    USE_EDGE_HIST: bool = True
    nPercentage: float = 0.3

    def nLevels(self) -> int:  # camerapyr.h:68-71
        return self.PYR_MIN_LVL - self.PYR_MAX_LVL + 1

    def _c_cfg(self) -> revo_pyr_config:
        return revo_pyr_config(self.nLevels(), self.cannyThreshold1, self.cannyThreshold2, self.DEPTH_MIN, self.DEPTH_MAX,
                               int(self.USE_EDGE_HIST), self.nPercentage, 20)

    def _c_cam(self) -> revo_camera:
        return revo_camera(self.fx, self.fy, self.cx, self.cy, self.width, self.height)


@dataclass
class Camera:
    """datastructures/camerapyr.h:90-111"""
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int

    @property
    def area(self) -> int:
        return self.width * self.height


class CameraPyr:
    """datastructures/camerapyr.h:113-193: levels 0..nLevels (one more than the pyramid uses)."""

    def __init__(self, settings: ImgPyramidSettings):
        self.camPyr: List[Camera] = []
        f32 = np.float32
        for lvl in range(settings.nLevels() + 1):
            if lvl == 0:
                self.camPyr.append(Camera(float(f32(settings.fx)), float(f32(settings.fy)), float(f32(settings.cx)),
                                          float(f32(settings.cy)), settings.width, settings.height))
            else:
                s = f32(1.0 / 2.0 ** lvl)
                self.camPyr.append(Camera(float(f32(settings.fx) * s), float(f32(settings.fy) * s), float(f32(settings.cx) * s),
                                          float(f32(settings.cy) * s), int(f32(settings.width) * s), int(f32(settings.height) * s)))

    def size(self) -> int:
        return len(self.camPyr)

    def at(self, lvl: int) -> Camera:
        return self.camPyr[lvl]


@dataclass
class OptimizerSettings:
    """system/optimizer.h:42-112."""
    lambdaSuccessFac: float = 0.5
    lambdaFailFac: float = 2.0
    lambdaInitial: List[float] = field(default_factory=lambda: [0.0] * 6)
    stepSizeMin: List[float] = field(default_factory=lambda: [1e-16] * 6)
    convergenceEps: List[float] = field(default_factory=lambda: [0.999] * 6)
    maxItsPerLvl: List[int] = field(default_factory=lambda: [100] * 6)
    edgeDistanceLvl: List[float] = field(default_factory=lambda: [30, 20, 10, 5, 5, 5])
    huber_edge: float = 0.3
    USE_EDGE_FILTER: bool = False   # OptimizerSettings() default, optimizer.h:80
    max_lm_tries: int = 0           # extension: fixed-iteration test mode (0 = reference behaviour)

    def _c(self) -> revo_opt_config:
        c = revo_opt_config()
        c.lambda_success_fac = self.lambdaSuccessFac
        c.lambda_fail_fac = self.lambdaFailFac
        for l in range(MAX_LEVELS):
            c.lambda_initial[l] = self.lambdaInitial[l]
            c.step_size_min[l] = self.stepSizeMin[l]
            c.convergence_eps[l] = self.convergenceEps[l]
            c.max_its_per_lvl[l] = self.maxItsPerLvl[l]
            c.edge_distance_lvl[l] = self.edgeDistanceLvl[l]
        c.huber_edge = self.huber_edge
        c.use_edge_filter = int(self.USE_EDGE_FILTER)
        c.max_lm_tries = self.max_lm_tries
        return c


@dataclass
class TrackerSettings:
    """system/tracker.h:31-50: USE_EDGE_FILTER defaults to true here (tracker.h:46)."""
    CHECK_INIT_VALUES: bool = True
    optimizerSettings: OptimizerSettings = field(default_factory=lambda: OptimizerSettings(USE_EDGE_FILTER=True))
    CHECK_TRACKING_RESULTS: bool = True     # tracker.h:45
    nFramesHistogramVoting: int = 3         # tracker.h:44,47


@dataclass
class ResidualInfo:
    """Optimizer::ResidualInfo, system/optimizer.h:117-139."""
    goodPtsEdges: int = 0
    badPtsEdges: int = 0
    sumErrorUnweighted: float = 0.0
    sumErrorWeighted: float = 0.0


# ---------------------------------------------------------------------------
# context
# ---------------------------------------------------------------------------
class Context:
    """Owns the device, stream and scratch memory (``revo_ctx``)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.revo_ctx_create(device, C.byref(h))
        if rc:
            raise RevoError(rc, self.lib.revo_strerror(rc).decode())
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc:
            msg = self.lib.revo_strerror(rc).decode()
            if rc == 3:
                msg += ": " + self.lib.revo_last_error(self.h).decode()
            raise RevoError(rc, msg)

    def synchronize(self):
        self.check(self.lib.revo_ctx_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.revo_ctx_stream(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.revo_ctx_launch_count(self.h))

    def last_timings(self):
        """(pyramid_ms, keyframe_ms, track_kernel_ms) of the most recent launches (CUDA events on the context stream)."""
        a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
        self.check(self.lib.revo_ctx_last_timings(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def last_upload_ms(self) -> float:
        a = C.c_float(0)
        self.check(self.lib.revo_ctx_last_upload_ms(self.h, C.byref(a)))
        return a.value

    def set_track_shape(self, ctas_per_pair: int = 0, threads_per_cta: int = 0):
        self.check(self.lib.revo_ctx_set_track_shape(self.h, ctas_per_pair, threads_per_cta))

    def set_track_max_clusters(self, max_clusters: int = 0):
        """Cap on the resident clusters of the tracking kernel (0 = all): room for the build kernels of a second context."""
        self.check(self.lib.revo_ctx_set_track_max_clusters(self.h, max_clusters))

    def set_track_engine(self, engine: int = 0, chunk_points: int = 0):
        """Kept for ABI compatibility: 0 / 1 = the cluster engine (the only one), anything else raises REVO_ERR_UNSUPPORTED."""
        self.check(self.lib.revo_ctx_set_track_engine(self.h, engine, chunk_points))

    def reserve(self, nbytes: int):
        """Pre-size the device memory pool (see revo_ctx_reserve)."""
        self.check(self.lib.revo_ctx_reserve(self.h, int(nbytes)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.revo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(a) -> int:
    """Host numpy array or an integer device pointer / object exposing data_ptr()."""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return int(a)


# ---------------------------------------------------------------------------
# ImgPyramidRGBD
# ---------------------------------------------------------------------------
def colored_pcl_from_arrays(bgr, depth, edges, fx, fy, cx, cy, depth_min, depth_max, dense: bool) -> np.ndarray:
    """The loop of ``ImgPyramidRGBD::generateColoredPcl`` (imgpyramidrgbd.cpp:300-323) vectorised: pixels in the reference's
    column-major scan order (x outer, y inner) with finite ``depth_min < Z < depth_max`` and -- unless ``dense`` -- an edge
    label; column = ``(Z(x-cx)/fx, Z(y-cy)/fy, Z, 1, R/255, G/255, B/255, 1)`` in float32 arithmetic."""
    depth = np.asarray(depth, np.float32)
    with np.errstate(invalid="ignore"):
        ok = np.isfinite(depth) & (depth > np.float32(depth_min)) & (depth < np.float32(depth_max))
    if not dense:
        ok &= np.asarray(edges) > 0
    xs, ys = np.nonzero(ok.T)                        # transposed: x-major order
    Z = depth[ys, xs]
    out = np.empty((8, len(xs)), np.float32)
    out[0] = Z * (xs.astype(np.float32) - np.float32(cx)) / np.float32(fx)
    out[1] = Z * (ys.astype(np.float32) - np.float32(cy)) / np.float32(fy)
    out[2] = Z
    out[3] = 1.0
    clr = np.asarray(bgr)[ys, xs].astype(np.float32) / np.float32(255.0)
    out[4], out[5], out[6] = clr[:, 2], clr[:, 1], clr[:, 0]
    out[7] = 1.0
    return out


class ImgPyramidRGBD:
    """``ImgPyramidRGBD(settings, cameraPyr, fullResRgb, fullResDepth, timestamp)`` --
    datastructures/imgpyramidrgbd.h:39-41.  ``rgb`` is HxWx3|4 uint8 in OpenCV BGR order,
    ``depth`` HxW float32 metres (0/NaN invalid).  Use :meth:`create_batch` for many frames."""

    def __init__(self, ctx: Context, settings: ImgPyramidSettings, cameraPyr: Optional[CameraPyr], rgb, depth,
                 timestamp: float = 0.0, _handle=None):
        self.ctx = ctx
        self.mSettings = settings
        self.cameraPyr = cameraPyr or CameraPyr(settings)
        self.frameId = 0
        self._T_w_f = np.eye(4, dtype=np.float32)
        self.rgbFullSize = None
        if _handle is not None:
            self.h = _handle
            return
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        self.rgbFullSize = rgb                      # imgpyramidrgbd.cpp:51 keeps a clone for generateColoredPcl (viewer)
        assert rgb.ndim == 3 and rgb.shape[2] in (3, 4) and rgb.shape[:2] == (settings.height, settings.width)
        assert depth.shape == (settings.height, settings.width)
        cfg, cam = settings._c_cfg(), settings._c_cam()
        h = C.c_void_p()
        ctx.check(ctx.lib.revo_pyr_create(ctx.h, C.byref(cfg), C.byref(cam), rgb.ctypes.data, 0, rgb.shape[2], depth.ctypes.data,
                                          0, timestamp, C.byref(h)))
        ctx.synchronize()   # rgb/depth temporaries may be released by the caller
        self.h = h

    @staticmethod
    def create_batch(ctx: Context, settings: ImgPyramidSettings, rgb, depth, timestamps: Optional[Sequence[float]] = None,
                     n: Optional[int] = None, channels: int = 3, cameraPyr: Optional[CameraPyr] = None,
                     synchronize: bool = True) -> List["ImgPyramidRGBD"]:
        """n frames in one set of batched launches. rgb: (n,H,W,C) uint8, depth: (n,H,W) float32 --
        host numpy arrays or device pointers / torch tensors (then pass ``n``)."""
        if isinstance(rgb, np.ndarray):
            rgb = np.ascontiguousarray(rgb, np.uint8)
            depth = np.ascontiguousarray(depth, np.float32)
            n = rgb.shape[0]
            channels = rgb.shape[3]
            assert depth.shape == rgb.shape[:3]
        assert n is not None
        cfg, cam = settings._c_cfg(), settings._c_cam()
        hs = (C.c_void_p * n)()
        ts = None
        if timestamps is not None:
            ts = np.ascontiguousarray(timestamps, np.float64)
        ctx.check(ctx.lib.revo_pyr_create_batch(ctx.h, C.byref(cfg), C.byref(cam), n, _ptr(rgb), channels, _ptr(depth),
                                                _ptr(ts), hs))
        if synchronize:
            ctx.synchronize()
        cp = cameraPyr or CameraPyr(settings)
        return [ImgPyramidRGBD(ctx, settings, cp, None, None, _handle=C.c_void_p(hs[i])) for i in range(n)]

    # -- keyframe ---------------------------------------------------------
    def makeKeyframe(self):
        """imgpyramidrgbd.cpp:231-252"""
        self.ctx.check(self.ctx.lib.revo_pyr_make_keyframe(self.ctx.h, self.h))

    @staticmethod
    def makeKeyframes(ctx: Context, pyrs: Sequence["ImgPyramidRGBD"]):
        arr = (C.c_void_p * len(pyrs))(*[p.h for p in pyrs])
        ctx.check(ctx.lib.revo_pyr_make_keyframe_batch(ctx.h, len(pyrs), arr))

    def isKeyframe(self) -> bool:
        return bool(self.ctx.lib.revo_pyr_is_keyframe(self.h))

    # -- accessors (imgpyramidrgbd.h:45-117) ---------------------------------
    def _cam(self, lvl) -> revo_camera:
        cam = revo_camera()
        self.ctx.check(self.ctx.lib.revo_pyr_level_camera(self.h, lvl, C.byref(cam)))
        return cam

    def _download(self, lvl: int, which: int, dtype, shape_fn):
        cam = self._cam(lvl)
        nbytes = C.c_size_t(0)
        self.ctx.check(self.ctx.lib.revo_pyr_download(self.ctx.h, self.h, lvl, which, None, 0, C.byref(nbytes)))
        out = np.empty(nbytes.value // np.dtype(dtype).itemsize, dtype)
        if nbytes.value:
            self.ctx.check(self.ctx.lib.revo_pyr_download(self.ctx.h, self.h, lvl, which, out.ctypes.data, out.nbytes, None))
        return out.reshape(shape_fn(cam))

    def returnGray(self, lvl):
        return self._download(lvl, GRAY, np.uint8, lambda c: (c.height, c.width))

    def returnDepth(self, lvl):
        return self._download(lvl, DEPTH, np.float32, lambda c: (c.height, c.width))

    def returnEdges(self, lvl):
        return self._download(lvl, EDGES, np.uint8, lambda c: (c.height, c.width))

    def returnOrigEdges(self, lvl):
        # imgpyramidrgbd.h:69-77
        if self.mSettings.USE_EDGE_HIST and lvl > self.mSettings.PYR_MAX_LVL:
            return self._download(lvl, EDGES_ORIG, np.uint8, lambda c: (c.height, c.width))
        return self.returnEdges(lvl)

    def returnHist(self, lvl):
        P = max(1, 20 >> lvl)
        return self._download(lvl, HIST, np.uint8, lambda c: (c.height // P, c.width // P))

    def return3DEdges(self, lvl):
        """(N,4) float32, rows in the reference's column-major scan order (the reference's 4xN
        Eigen::MatrixXf is column-major, i.e. the same memory)."""
        return self._download(lvl, EDGES3D, np.float32, lambda c: (-1, 4))

    def return3DEdgesDeviceOrder(self, lvl):
        return self._download(lvl, EDGES3D_DEVICE_ORDER, np.float32, lambda c: (-1, 4))

    def returnDistTransform(self, lvl):
        return self._download(lvl, DT, np.float32, lambda c: (c.height, c.width))

    def returnOptimizationStructure(self, lvl):
        """Raises RevoError(REVO_ERR_NOT_KEYFRAME) where the reference exit(0)s (imgpyramidrgbd.h:113-117)."""
        return self._download(lvl, OPTSTRUCT, np.float32, lambda c: (c.height, c.width, 4))

    def returnK(self, lvl):
        c = self._cam(lvl)
        return np.array([[c.fx, 0, c.cx], [0, c.fy, c.cy], [0, 0, 1]], np.float32)

    def returnNumEdges(self, lvl) -> int:
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.revo_pyr_num_edges(self.ctx.h, self.h, lvl, C.byref(n)))
        return n.value

    def returnTimestamp(self) -> float:
        return float(self.ctx.lib.revo_pyr_timestamp(self.h))

    def returnMinLvl(self):
        return self.mSettings.PYR_MIN_LVL

    def returnMaxLvl(self):
        return self.mSettings.PYR_MAX_LVL

    def setTwf(self, T):
        self._T_w_f = np.asarray(T, np.float32).reshape(4, 4).copy()

    def getTransKFtoWorld(self):
        return self._T_w_f

    def isPointOkDepth(self, z) -> bool:
        return bool(np.isfinite(z) and self.mSettings.DEPTH_MIN < z < self.mSettings.DEPTH_MAX)

    def generateColoredPcl(self, lvl: int, densePcl: bool = False, rgb=None) -> np.ndarray:
        """``generateColoredPcl(lvl, clrPcl, densePcl)`` (imgpyramidrgbd.cpp:279-327): the viewer's coloured cloud, an (8, N)
        float32 matrix of columns ``(X, Y, Z, 1, r, g, b, 1)`` in the reference's column-major scan order, compacted on the
        device (``revo_pyr_colored_pcl``).  ``rgb``: the full-resolution colour image (default: the one kept at construction,
        like the reference's ``rgbFullSize`` clone; pyramids taken from a batch must pass it)."""
        rgb = self.rgbFullSize if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        if rgb is None:
            raise RevoError(REVO_ERR_INVALID_ARG, "generateColoredPcl needs the colour image (pyramid built from a batch handle)")
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.revo_pyr_colored_pcl(self.ctx.h, self.h, lvl, int(densePcl), rgb.ctypes.data, rgb.shape[2], None, 0,
                                                         C.byref(n)))
        out = np.zeros((n.value, 8), np.float32)
        if n.value:
            self.ctx.check(self.ctx.lib.revo_pyr_colored_pcl(self.ctx.h, self.h, lvl, int(densePcl), rgb.ctypes.data, rgb.shape[2],
                                                             out.ctypes.data, n.value, C.byref(n)))
        return out.T.copy()

    # -- test hook -----------------------------------------------------------
    def uploadLevel(self, lvl, pts4=None, dt=None, opt4=None):
        n = 0
        if pts4 is not None:
            pts4 = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
            n = len(pts4)
        if dt is not None:
            dt = np.ascontiguousarray(dt, np.float32)
        if opt4 is not None:
            opt4 = np.ascontiguousarray(opt4, np.float32)
        self.ctx.check(self.ctx.lib.revo_pyr_upload_level(self.ctx.h, self.h, lvl, _ptr(pts4), n, _ptr(dt), _ptr(opt4)))

    def destroy(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owned", True):
            self.ctx.lib.revo_pyr_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class PyramidBatch:
    """n ImgPyramidRGBD handles created by ONE revo_pyr_create_batch call, kept as a raw handle array: what the
    throughput path (bench.py, stream.py) passes around instead of n python objects.  ``batch[i]`` gives a
    non-owning :class:`ImgPyramidRGBD` view for the accessors."""

    __slots__ = ("ctx", "settings", "n", "arr", "campyr", "_alive")

    def __init__(self, ctx: Context, settings: ImgPyramidSettings, rgb, depth, n: int, channels: int = 3, timestamps=None,
                 cameraPyr: Optional[CameraPyr] = None, depth_scale_factor: float = 5000.0):
        """depth: float32 metres, or uint16 raw sensor / dataset values (then metres = raw / depth_scale_factor, converted
        on the device like the reference's reader does on the host, io/iowrapperRGBD.cpp:327)."""
        self.ctx, self.settings, self.n, self.campyr = ctx, settings, n, cameraPyr
        cfg, cam = settings._c_cfg(), settings._c_cam()
        self.arr = (C.c_void_p * n)()
        ts = None if timestamps is None else np.ascontiguousarray(timestamps, np.float64)
        if "int16" in str(getattr(depth, "dtype", "")):     # uint16 (numpy / torch), or int16 holding the same bits
            scale = float(np.float32(1.0) / np.float32(depth_scale_factor))
            ctx.check(ctx.lib.revo_pyr_create_batch_u16(ctx.h, C.byref(cfg), C.byref(cam), n, _ptr(rgb), channels, _ptr(depth),
                                                        scale, _ptr(ts), self.arr))
        else:
            ctx.check(ctx.lib.revo_pyr_create_batch(ctx.h, C.byref(cfg), C.byref(cam), n, _ptr(rgb), channels, _ptr(depth), _ptr(ts),
                                                    self.arr))
        self._alive = True

    def __len__(self):
        return self.n

    def __getitem__(self, i) -> "ImgPyramidRGBD":
        v = ImgPyramidRGBD(self.ctx, self.settings, self.campyr, None, None, _handle=C.c_void_p(self.arr[i]))
        v._owned = False
        return v

    def makeKeyframes(self, ctx: Optional[Context] = None):
        ctx = ctx or self.ctx
        ctx.check(ctx.lib.revo_pyr_make_keyframe_batch(ctx.h, self.n, self.arr))

    def destroy(self, ctx: Optional[Context] = None):
        """Release the batch.  ctx: the context whose stream orders the release (default: the creating context).  A
        pipeline that builds on one context and tracks on another should release on the TRACKING context: its stream is
        past the last use, so the memory is immediately reusable by the upload of a later batch."""
        ctx = ctx or self.ctx
        if self._alive and getattr(ctx, "h", None):
            ctx.lib.revo_pyr_destroy_batch(ctx.h, self.n, self.arr)
        self._alive = False

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _handles(frames):
    """(n, ctypes handle array) of a PyramidBatch or a sequence of ImgPyramidRGBD."""
    if isinstance(frames, PyramidBatch):
        return frames.n, frames.arr
    if isinstance(frames, np.ndarray):        # raw handle values (uint64), see stream.CudaBackend.take
        assert frames.dtype == np.uint64 and frames.flags.c_contiguous
        return int(frames.size), frames.ctypes.data_as(C.POINTER(C.c_void_p))
    n = len(frames)
    return n, (C.c_void_p * n)(*[p.h for p in frames])


def quat_to_R(q_xyzw) -> np.ndarray:
    """Unit quaternion (x, y, z, w; normalised by the library) -> 3x3 rotation (``revo_quat_to_R9``; host arithmetic)."""
    lib = load_library()
    q = np.ascontiguousarray(q_xyzw, np.float32).reshape(4)
    r9 = np.zeros(9, np.float32)
    rc = lib.revo_quat_to_R9(q.ctypes.data, r9.ctypes.data)
    if rc:
        raise RevoError(rc, lib.revo_strerror(rc).decode())
    return _R_from_c(r9)


def R_to_quat(R) -> np.ndarray:
    """3x3 rotation -> (x, y, z, w) like ``Sophus::SO3f(R).unit_quaternion()``; raises RevoError(REVO_ERR_NOT_ORTHOGONAL)
    where Sophus would abort (``revo_R9_to_quat``; host arithmetic)."""
    lib = load_library()
    r9 = _R_to_c(R)
    q = np.zeros(4, np.float32)
    rc = lib.revo_R9_to_quat(r9.ctypes.data, q.ctypes.data)
    if rc:
        raise RevoError(rc, lib.revo_strerror(rc).decode())
    return q


def _R_to_c(R) -> np.ndarray:
    """3x3 numpy (row-major view of the matrix) -> Eigen column-major 9 floats."""
    return np.ascontiguousarray(np.asarray(R, np.float32).reshape(3, 3).T.reshape(-1))


def _R_from_c(r9) -> np.ndarray:
    return np.asarray(r9, np.float32).reshape(3, 3).T.copy()


# ---------------------------------------------------------------------------
# Optimizer
# ---------------------------------------------------------------------------
class Optimizer:
    """system/optimizer.h:114-186."""

    def __init__(self, ctx: Context, settings: Optional[OptimizerSettings] = None):
        self.ctx = ctx
        self.mSettings = settings or OptimizerSettings()

    def trackFrames(self, refFrame: ImgPyramidRGBD, currFrame: ImgPyramidRGBD, R, T, lvl: int, resInfo: ResidualInfo):
        """``float Optimizer::trackFrames(ref, cur, R&, T&, lvl, resInfo&)`` (optimizer.cpp:235-311).
        Returns (error, R, T); resInfo is updated in place; ``self.last_n_evals`` holds the evaluation count."""
        cfg = self.mSettings._c()
        Rc, Tc = _R_to_c(R), np.ascontiguousarray(T, np.float32).copy()
        res = revo_residual_info()
        err = C.c_float(0)
        ne = C.c_int(0)
        self.ctx.check(self.ctx.lib.revo_track_level(self.ctx.h, C.byref(cfg), refFrame.h, currFrame.h, lvl, Rc.ctypes.data,
                                                     Tc.ctypes.data, C.byref(res), C.byref(err), C.byref(ne)))
        resInfo.goodPtsEdges, resInfo.badPtsEdges = res.good_pts_edges, res.bad_pts_edges
        resInfo.sumErrorUnweighted, resInfo.sumErrorWeighted = res.sum_error_unweighted, res.sum_error_weighted
        self.last_n_evals = ne.value
        return float(err.value), _R_from_c(Rc), Tc

    def evalRecord(self, refFrame, currFrame, R, T, lvl: int) -> np.ndarray:
        """One fused PASS A + PASS B evaluation: the 32-value record (see revo_eval)."""
        cfg = self.mSettings._c()
        Rc, Tc = _R_to_c(R), np.ascontiguousarray(T, np.float32)
        rec = np.zeros(32, np.float64)
        self.ctx.check(self.ctx.lib.revo_eval(self.ctx.h, C.byref(cfg), refFrame.h, currFrame.h, lvl, Rc.ctypes.data,
                                              Tc.ctypes.data, rec.ctypes.data))
        return rec


# ---------------------------------------------------------------------------
# TrackerNew
# ---------------------------------------------------------------------------
class TrackerNew:
    """system/tracker.h:52-112 (trackFrames + checkInitializationValues + evalCostFunction)."""

    def __init__(self, ctx: Context, config: Optional[TrackerSettings] = None, pyrConfig: Optional[ImgPyramidSettings] = None):
        self.ctx = ctx
        self.mSettings = config or TrackerSettings()
        self.mPyrConfig = pyrConfig or ImgPyramidSettings()
        self.histogramLevel = 2
        # past frames for the tracking-quality vote: (pyramid whose level-histogramLevel 3-D edge list is used, world pose, ts)
        self.mPastPcl: list = []
        self.last_quality: Optional[revo_quality_result] = None

    def _c_cfg(self) -> revo_tracker_config:
        c = revo_tracker_config()
        c.check_init_values = int(self.mSettings.CHECK_INIT_VALUES)
        c.pyr_min_lvl = self.mPyrConfig.PYR_MIN_LVL
        c.pyr_max_lvl = self.mPyrConfig.PYR_MAX_LVL
        c.opt = self.mSettings.optimizerSettings._c()
        return c

    def trackFrames(self, R, T, refFrame: ImgPyramidRGBD, currFrame: ImgPyramidRGBD):
        """``TrackerStatus trackFrames(R&, T&, error&, ref, cur)`` (tracker.cpp:294-353).
        Returns (status, R, T, error); the full result struct is kept in ``self.last_result``."""
        cfg = self._c_cfg()
        Rc, Tc = _R_to_c(R), np.ascontiguousarray(T, np.float32).copy()
        res = revo_track_result()
        self.ctx.check(self.ctx.lib.revo_track(self.ctx.h, C.byref(cfg), refFrame.h, currFrame.h, Rc.ctypes.data, Tc.ctypes.data,
                                               C.byref(res)))
        self.last_result = res
        return int(res.status), _R_from_c(Rc), Tc, float(res.error)

    def trackFramesBatch(self, Rs, Ts, refFrames: Sequence[ImgPyramidRGBD], currFrames: Sequence[ImgPyramidRGBD],
                         trace_cap: int = 0, check: bool = True):
        """n independent pairs in one persistent-kernel launch.  Rs: (n,3,3), Ts: (n,3).
        Returns a structured array (TRACK_RESULT_DTYPE; R column-major) and, if trace_cap>0, the LM traces.
        ``revo_track_batch`` answers REVO_OK when the launch worked even if single pairs were refused (their ``rc`` field says
        why, e.g. REVO_ERR_NOT_ORTHOGONAL for a bad initial rotation; such a pair comes back untracked).  With ``check`` (the
        default) a refused pair raises :class:`RevoError` naming it, like the single-pair call does; pass ``check=False`` to
        inspect ``out["rc"]`` yourself."""
        n, refs = _handles(refFrames)
        n2, curs = _handles(currFrames)
        assert n == n2
        cfg = self._c_cfg()
        Rc = np.ascontiguousarray(np.asarray(Rs, np.float32).reshape(n, 3, 3).transpose(0, 2, 1).reshape(n, 9))
        Tc = np.ascontiguousarray(np.asarray(Ts, np.float32).reshape(n, 3))
        out = np.zeros(n, TRACK_RESULT_DTYPE)
        trace = counts = None
        if trace_cap > 0:
            trace = (revo_trace_entry * (trace_cap * n))()
            counts = np.zeros(n, np.int32)
        self.ctx.check(self.ctx.lib.revo_track_batch(self.ctx.h, C.byref(cfg), n, refs, curs, Rc.ctypes.data, Tc.ctypes.data,
                                                     out.ctypes.data, C.addressof(trace) if trace is not None else None,
                                                     trace_cap, _ptr(counts)))
        if check and out["rc"].any():
            i = int(np.nonzero(out["rc"])[0][0])
            rc = int(out["rc"][i])
            raise RevoError(rc, f"pair {i} of {n} was not tracked: {self.ctx.lib.revo_strerror(rc).decode()} "
                                f"({int((out['rc'] != 0).sum())} pairs refused in this batch)")
        if trace_cap > 0:
            traces = [[(trace[i * trace_cap + k].error, trace[i * trace_cap + k].lam, trace[i * trace_cap + k].accepted,
                        trace[i * trace_cap + k].good, trace[i * trace_cap + k].bad, trace[i * trace_cap + k].level)
                       for k in range(counts[i])] for i in range(n)]
            return out, traces
        return out

    # -- tracking-quality vote (tracker.cpp:118-257) -----------------------------------
    def addOldPclAndPose(self, pyr, worldPose, timeStamp: float = 0.0):
        """``addOldPclAndPose(pcl, worldPose, ts)`` (tracker.cpp:209-224).  Like the reference, which stores
        ``return3DEdges(histogramLevel)`` by value, the tracker keeps a COPY of that one list (a :class:`PointList` on the
        device), not the pyramid: the frame and the batch it belongs to can be released.  ``pyr`` may also be a
        :class:`PointList` made earlier (``copy_point_lists`` copies the lists of many streams in one call)."""
        pl = pyr if isinstance(pyr, PointList) else copy_point_lists(self.ctx, [pyr], self.histogramLevel)[0]
        self.mPastPcl.append((pl, np.asarray(worldPose, np.float32).reshape(4, 4).copy(), float(timeStamp)))
        # The reference's lists grow until the next keyframe (the pop in addOldPclAndPose is commented out there), but only the
        # FIRST nFramesHistogramVoting entries ever vote (tracker.cpp:138) and clearUpPastLists keeps the LAST ones: entries
        # in between can never be read again and are dropped here.
        nv = self.mSettings.nFramesHistogramVoting
        while len(self.mPastPcl) > 2 * nv:
            self.mPastPcl.pop(nv)[0].destroy()

    def clearUpPastLists(self):
        """tracker.cpp:249-257"""
        while len(self.mPastPcl) > self.mSettings.nFramesHistogramVoting:
            self.mPastPcl.pop(0)[0].destroy()

    def _vote_inputs(self, estimatedPose):
        """The past lists / world poses that vote (column-major poses, three slots) and the estimated pose."""
        nv = min(len(self.mPastPcl), self.mSettings.nFramesHistogramVoting, 3)
        hs = [self.mPastPcl[f][0].h for f in range(nv)] + [None] * (3 - nv)
        poses = np.zeros((3, 16), np.float32)
        for f in range(nv):
            poses[f] = self.mPastPcl[f][1].T.reshape(-1)
        est = np.asarray(estimatedPose, np.float32).reshape(4, 4).T.reshape(-1)
        return len(self.mPastPcl), hs, poses, est

    def assessTrackingQuality(self, estimatedPose, currFrame: ImgPyramidRGBD) -> int:
        """``TrackerStatus assessTrackingQuality(estimatedPose, currFrame)`` (tracker.cpp:118-201); the counts are kept in
        ``self.last_quality``."""
        return assess_tracking_quality_batch([self], [estimatedPose], [currFrame])[0]

    # -- multi-GPU split (one process per GPU) ------------------------------------
    def splitExport(self, rank: int, world: int) -> bytes:
        buf = C.create_string_buffer(SPLIT_HANDLE_BYTES)
        self.ctx.check(self.ctx.lib.revo_split_export(self.ctx.h, rank, world, buf))
        return buf.raw

    def splitOpen(self, blobs: bytes):
        self.ctx.check(self.ctx.lib.revo_split_open(self.ctx.h, blobs))

    def trackFramesSplit(self, R, T, refFrame, currFrame):
        cfg = self._c_cfg()
        Rc, Tc = _R_to_c(R), np.ascontiguousarray(T, np.float32).copy()
        res = revo_track_result()
        self.ctx.check(self.ctx.lib.revo_track_split(self.ctx.h, C.byref(cfg), refFrame.h, currFrame.h, Rc.ctypes.data,
                                                     Tc.ctypes.data, C.byref(res)))
        self.last_result = res
        return int(res.status), _R_from_c(Rc), Tc, float(res.error)


class PointList:
    """A device copy of one level's 3-D edge list (what ``TrackerNew::addOldPclAndPose`` stores, tracker.cpp:209-224)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle

    def download(self, lvl: int) -> np.ndarray:
        """(N, 4) float32 in device (tile-major) order."""
        n = C.c_size_t(0)
        self.ctx.check(self.ctx.lib.revo_pyr_download(self.ctx.h, self.h, lvl, EDGES3D_DEVICE_ORDER, None, 0, C.byref(n)))
        out = np.zeros((n.value // 16, 4), np.float32)
        if n.value:
            self.ctx.check(self.ctx.lib.revo_pyr_download(self.ctx.h, self.h, lvl, EDGES3D_DEVICE_ORDER, out.ctypes.data, n.value, None))
        return out

    def destroy(self):
        if self.h:
            self.ctx.lib.revo_pyr_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def copy_point_lists(ctx: Context, pyrs, lvl: int) -> List["PointList"]:
    """``revo_pyr_copy_points_batch``: device copies of the level-``lvl`` 3-D edge lists of many pyramids (one allocation, one
    launch)."""
    n, hs = _handles(pyrs)
    out = (C.c_void_p * n)()
    ctx.check(ctx.lib.revo_pyr_copy_points_batch(ctx.h, n, hs, lvl, out))
    return [PointList(ctx, out[i]) for i in range(n)]


def assess_tracking_quality_batch(trackers: Sequence["TrackerNew"], estimatedPoses, currFrames) -> List[int]:
    """``assessTrackingQuality`` of many streams in one launch pair (``revo_track_quality_batch``): tracker ``i`` votes on
    ``currFrames[i]`` under ``estimatedPoses[i]`` with its own history.  Returns the statuses; every tracker's
    ``last_quality`` is set like the single call does."""
    n = len(trackers)
    status = [TRACKER_STATE_OK] * n
    todo = [i for i in range(n) if trackers[i].mPastPcl and trackers[i].mSettings.CHECK_TRACKING_RESULTS]
    if not todo:
        return status
    t0 = trackers[todo[0]]
    ctx = t0.ctx
    m = len(todo)
    curs = (C.c_void_p * m)(*[currFrames[i].h for i in todo])
    past = (C.c_void_p * (3 * m))()
    n_past = np.zeros(m, np.int32)
    poses = np.zeros((m, 3, 16), np.float32)
    est = np.zeros((m, 16), np.float32)
    for k, i in enumerate(todo):
        n_past[k], hs, poses[k], est[k] = trackers[i]._vote_inputs(estimatedPoses[i])
        for f in range(3):
            past[3 * k + f] = hs[f]
    res = (revo_quality_result * m)()
    ctx.check(ctx.lib.revo_track_quality_batch(ctx.h, m, curs, t0.histogramLevel, n_past.ctypes.data, past, poses.ctypes.data,
                                               est.ctypes.data, t0.mSettings.nFramesHistogramVoting, res))
    for k, i in enumerate(todo):
        r = revo_quality_result()
        C.memmove(C.byref(r), C.byref(res[k]), C.sizeof(revo_quality_result))
        trackers[i].last_quality = r
        status[i] = int(r.status)
    return status


def result_R(res_row) -> np.ndarray:
    """Row of TRACK_RESULT_DTYPE -> 3x3 numpy matrix."""
    return _R_from_c(res_row["R"])
