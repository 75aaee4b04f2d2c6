"""Minimal multi-stream driver of the hot path: B independent RGB-D streams advance one frame per step.

This is the slice of ``REVO::start`` (system/system.cpp:128-283) the tracker needs to run on a stream:
the motion-model initialisation ``T_init = T_kf_N * T_NM1_N`` (system.cpp:262-271), world-pose
composition, and promotion of a frame to keyframe (``makeKeyframe`` + re-initialisation with
``T_NM1_N``, system.cpp:203-216).  The tracking-quality vote that decides WHEN to switch keyframes
(``assessTrackingQuality``) is a "next" row; here a keyframe is promoted every
``kf_interval`` frames.

The backend object does the actual work (the CUDA library in production; bench.py plugs the CPU oracle
in for the reference arm) and must provide::

    create(bgr, depth, n) -> list of frame handles        (ImgPyramidRGBD construction)
    make_keyframes(handles)                                (makeKeyframe)
    track(Rs, Ts, refs, curs) -> dict(R (n,3,3), T (n,3), status (n,), n_evals (n,6), n_pts (n,6))
    destroy(handles)
"""
from __future__ import annotations

import numpy as np


def _inv(T: np.ndarray) -> np.ndarray:
    """Batched inverse of rigid 4x4 transforms."""
    R = T[:, :3, :3]
    t = T[:, :3, 3]
    Ti = np.tile(np.eye(4, dtype=T.dtype), (T.shape[0], 1, 1))
    Rt = R.transpose(0, 2, 1)
    Ti[:, :3, :3] = Rt
    Ti[:, :3, 3] = -np.einsum("nij,nj->ni", Rt, t)
    return Ti


class StreamTracker:
    def __init__(self, backend, n_streams: int, kf_interval: int = 10):
        self.be = backend
        self.B = n_streams
        self.kf_interval = kf_interval
        self.kf = None                      # keyframe handles (one per stream)
        self.prev = None                    # previous-frame handles
        eye = np.tile(np.eye(4, dtype=np.float32), (n_streams, 1, 1))
        self.T_w_kf = eye.copy()            # keyframe pose in the world
        self.T_kf_prev = eye.copy()         # previous frame relative to its keyframe (T_kf_N)
        self.T_nm1_n = eye.copy()           # last inter-frame motion (T_NM1_N)
        self.T_w_c = eye.copy()             # current world pose of every stream
        self.frame = 0
        self.total_evals = 0
        self.total_point_evals = 0          # sum over pairs/levels of n_pts * n_evals (roofline numerator / 60 B)
        self.last = None
        self.keep_history = False           # bench.py parity block: per-step world poses and evaluation counts
        self.history = []
        self._pending = []                  # batches whose upload + build are in flight, oldest first

    def start(self, bgr, depth):
        """First frame of every stream: becomes the keyframe (system.cpp:151-175)."""
        if hasattr(self.be, "reserve"):
            self.be.reserve(self.B)          # size the device memory pool for the steady state before the first frame
        self.kf = self.be.create(bgr, depth, self.B)
        self.be.wait_created()
        self.be.make_keyframes(self.kf)
        self.prev = None
        self.frame = 0

    def prefetch(self, bgr, depth):
        """Start uploading / building the pyramids of a FUTURE frame: everything is enqueued asynchronously (upload on the
        backend's copy stream, kernels on its build stream), so it overlaps the tracking of earlier frames.  Call it
        twice before the first step_pipelined() to keep two frames in flight (upload of k+2 | build of k+1 | track of k)."""
        self._pending.append(self.be.create(bgr, depth, self.B))

    def step_pipelined(self, next_bgr=None, next_depth=None):
        """Enqueue the upload + build of one more future frame, then track the oldest frame in flight.  No host wait for
        the build: the tracking stream waits for the batch's build-complete event on the device."""
        if next_bgr is not None:
            self._pending.append(self.be.create(next_bgr, next_depth, self.B))
        return self._track(self._pending.pop(0))

    def step(self, bgr, depth):
        """Build the pyramids of the next frame of every stream and track it against its keyframe."""
        return self._track(self.be.create(bgr, depth, self.B))

    def _track(self, cur):
        T_init = self.T_kf_prev @ self.T_nm1_n                     # system.cpp:268
        out = self.be.track(T_init[:, :3, :3], T_init[:, :3, 3], self.kf, cur)
        T_kf_n = np.tile(np.eye(4, dtype=np.float32), (self.B, 1, 1))
        T_kf_n[:, :3, :3] = out["R"]
        T_kf_n[:, :3, 3] = out["T"]
        self.T_nm1_n = _inv(self.T_kf_prev) @ T_kf_n               # system.cpp:266
        self.T_w_c = self.T_w_kf @ T_kf_n                          # system.cpp:192
        self.frame += 1
        self.total_evals += int(out["n_evals"].sum())
        self.total_point_evals += int((out["n_evals"].astype(np.int64) * out["n_pts"].astype(np.int64)).sum())
        self.last = out
        if self.keep_history:
            self.history.append((self.T_w_c.copy(), T_kf_n.copy(), np.asarray(out["n_evals"]).copy()))
        if self.prev is not None:
            self.be.destroy(self.prev)
        if self.frame % self.kf_interval == 0:
            # promote the frame just tracked (the reference promotes prevPyr and re-tracks; with a fixed
            # interval the promoted frame's pose is already known): system.cpp:203-216
            self.be.make_keyframes(cur)
            self.be.destroy(self.kf)
            self.kf = cur
            self.prev = None
            self.T_w_kf = self.T_w_c.copy()
            self.T_kf_prev = np.tile(np.eye(4, dtype=np.float32), (self.B, 1, 1))
        else:
            self.prev = cur
            self.T_kf_prev = T_kf_n
        return out

    def close(self):
        for h in [self.prev, self.kf] + list(self._pending):
            if h is not None:
                self.be.destroy(h)
        self.prev = self.kf = None
        self._pending = []


class CudaBackend:
    """The product path: everything through the C ABI (revo_b200/api.py)."""

    def __init__(self, ctx, settings, tracker_settings=None, build_ctx=None):
        """ctx: context (stream) that tracks and promotes keyframes; build_ctx: optional second context whose stream
        uploads frames and builds pyramids, so that the H2D copy of frame k+1 overlaps the tracking of frame k."""
        from . import api

        self.api = api
        self.ctx = ctx
        self.build_ctx = build_ctx or ctx
        self.settings = settings
        self.tracker = api.TrackerNew(ctx, tracker_settings or api.TrackerSettings(), settings)
        self.campyr = api.CameraPyr(settings)

    def reserve(self, n_streams: int):
        """Steady state of a stream batch: keyframe + previous + current (+ one being built) frame slabs and two
        generations of keyframe structures (the new one is built before the old one is released)."""
        px = sum((self.settings.width >> l) * (self.settings.height >> l) for l in range(self.settings.nLevels()))
        frame = px * (1 + 4 + 1 + 1 + 8) + self.settings.width * self.settings.height * 5 + 65536    # gray, depth, edges x2, list; labels + counters
        kf = px * 36 + 4096
        total = int(n_streams * (5 * frame + 2 * kf + self.settings.width * self.settings.height * 3) * 1.1) + (64 << 20)
        self.ctx.reserve(total)

    def create(self, bgr, depth, n):
        return self.api.PyramidBatch(self.build_ctx, self.settings, bgr, depth, n, channels=3, cameraPyr=self.campyr)

    def wait_created(self):
        if self.build_ctx is not self.ctx:
            self.build_ctx.synchronize()

    def make_keyframes(self, handles):
        handles.makeKeyframes(self.ctx)

    def track(self, Rs, Ts, refs, curs):
        out = self.tracker.trackFramesBatch(Rs, Ts, refs, curs)
        n = len(refs)
        R = out["R"].reshape(n, 3, 3).transpose(0, 2, 1)   # column-major 9 floats -> (n, 3, 3)
        return dict(R=R, T=out["t"], status=out["status"], n_evals=out["n_evals"], n_pts=out["n_pts"], error=out["error"])

    def destroy(self, handles):
        handles.destroy(self.ctx)     # on the tracking stream (past the last use), not behind the builds in flight
