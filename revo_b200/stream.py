"""Minimal multi-stream driver of the hot path: B independent RGB-D streams advance one frame per step.

This is the slice of ``REVO::start`` (system/system.cpp:128-283) the tracker needs to run on a stream:
the motion-model initialisation ``T_init = T_kf_N * T_NM1_N`` (system.cpp:262-271), world-pose
composition, and promotion of a frame to keyframe (``makeKeyframe`` + re-initialisation with
``T_NM1_N``, system.cpp:203-216).  Two keyframe policies:

* ``kf_policy="interval"``: a keyframe every ``kf_interval`` frames (the fixed workload BASELINE.json's configs[1] is timed on);
* ``kf_policy="vote"``: the reference's own policy (system.cpp:199-239) -- after every alignment the tracking-quality vote
  (``assessTrackingQuality``, tracker.cpp:118-201) runs for ALL streams in one launch pair; the streams whose vote says
  NEW_KF (and that did not just switch) promote their PREVIOUS frame, are aligned again against it starting from the last
  inter-frame motion, and vote again -- one more (smaller) launch of each kind per step.  Vectorised over the streams; the
  per-stream object version of the same logic is :class:`revo_b200.system.MultiStreamREVO` (the tests tie the two together).

The backend object does the actual work (the CUDA library in production; bench.py plugs the CPU oracle
in for the reference arm) and must provide::

    create(bgr, depth, n) -> list of frame handles        (ImgPyramidRGBD construction)
    make_keyframes(handles)                                (makeKeyframe)
    track(Rs, Ts, refs, curs) -> dict(R (n,3,3), T (n,3), status (n,), n_evals (n,6), n_pts (n,6))
    destroy(handles)

and, for the vote policy, the handle-array flavour (``uint64`` numpy arrays of frame handles)::

    take(handles) -> uint64 array                          (the batch object gives up ownership of its frames)
    make_keyframes_h(h), destroy_h(h), track_h(Rs, Ts, refs_h, curs_h)
    copy_points_h(h, lvl) -> uint64 array                  (addOldPclAndPose's copy of return3DEdges(histogramLevel))
    vote_h(curs_h, n_past (n,), past_h (n,3), past_poses (n,3,4,4), est (n,4,4)) -> status (n,)
"""
from __future__ import annotations

import numpy as np


def _inv(T: np.ndarray) -> np.ndarray:
    """Batched inverse of 4x4 transforms: the general inverse like the reference's ``Matrix4f::inverse()`` (system.h:136-139), in
    float64 and rounded once, the same as revo_b200.system._inv."""
    return np.linalg.inv(T.astype(np.float64)).astype(np.float32)


TRACKER_STATE_OK, TRACKER_STATE_LOST, TRACKER_STATE_NEW_KF = 0, 1, 2
PIPELINED_TRACK_CLUSTERS = 62    # of 74 on B200 (see CudaBackend)
HISTOGRAM_LEVEL = 2        # TrackerNew::histogramLevel (tracker.cpp:229)
N_VOTING = 3               # TrackerSettings::nFramesHistogramVoting


class StreamTracker:
    def __init__(self, backend, n_streams: int, kf_interval: int = 10, kf_policy: str = "interval"):
        assert kf_policy in ("interval", "vote")
        self.be = backend
        self.B = n_streams
        self.kf_interval = kf_interval
        self.kf_policy = kf_policy
        self.n_keyframes = 0                # promotions after the first frame, all streams
        self.n_retracks = 0
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
        if self.kf_policy == "vote":
            B = self.B
            self.kf_h = self.be.take(self.kf)                     # per-stream keyframe handles
            self.kf = None
            self.prev_h = self.kf_h.copy()                        # system.cpp:153: kfPyr = prevPyr = currPyr
            self.T_w_prev = self.T_w_c.copy()
            self.just_added = np.ones(B, bool)
            # vote history per stream: up to 2 * N_VOTING lists (the first N_VOTING vote, the last N_VOTING survive a clear-up)
            self.past_h = np.zeros((B, 2 * N_VOTING), np.uint64)
            self.past_T = np.tile(np.eye(4, dtype=np.float32), (B, 2 * N_VOTING, 1, 1))
            self.n_past = np.zeros(B, np.int64)
            self._add_old(np.arange(B), self.kf_h, self.T_w_c)    # system.cpp:174

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

    # ---- the reference's keyframe policy, vectorised over the streams ------------------------------------------------
    def _add_old(self, idx, frames_h, T_w):
        """addOldPclAndPose (tracker.cpp:209-224) for the streams idx: one batched copy of the level-2 lists."""
        lists = self.be.copy_points_h(frames_h, HISTOGRAM_LEVEL)
        full = self.n_past[idx] >= 2 * N_VOTING
        if full.any():                                            # drop entry N_VOTING: it can never vote nor survive a clear-up
            f = idx[full]
            self.be.destroy_h(self.past_h[f, N_VOTING].copy())
            self.past_h[f, N_VOTING:-1] = self.past_h[f, N_VOTING + 1:]
            self.past_T[f, N_VOTING:-1] = self.past_T[f, N_VOTING + 1:]
            self.n_past[f] -= 1
        self.past_h[idx, self.n_past[idx]] = lists
        self.past_T[idx, self.n_past[idx]] = T_w[idx] if T_w.shape[0] == self.B else T_w
        self.n_past[idx] += 1

    def _clear_up(self, idx):
        """clearUpPastLists (tracker.cpp:249-257): keep the last N_VOTING entries."""
        for i in idx:
            n = int(self.n_past[i])
            if n > N_VOTING:
                self.be.destroy_h(self.past_h[i, :n - N_VOTING].copy())
                self.past_h[i, :N_VOTING] = self.past_h[i, n - N_VOTING:n]
                self.past_T[i, :N_VOTING] = self.past_T[i, n - N_VOTING:n]
                self.past_h[i, N_VOTING:] = 0
                self.n_past[i] = N_VOTING

    def _vote(self, idx, cur_h, T_w_c):
        return self.be.vote_h(cur_h[idx], self.n_past[idx], self.past_h[idx, :N_VOTING], self.past_T[idx, :N_VOTING], T_w_c[idx])

    def _track_vote(self, cur):
        B, eye = self.B, np.tile(np.eye(4, dtype=np.float32), (self.B, 1, 1))
        cur_h = self.be.take(cur)
        T_init = self.T_kf_prev @ self.T_nm1_n                     # system.cpp:268
        out = self.be.track_h(T_init[:, :3, :3], T_init[:, :3, 3], self.kf_h, cur_h)
        T_kf_n = eye.copy()
        T_kf_n[:, :3, :3] = out["R"]
        T_kf_n[:, :3, 3] = out["T"]
        T_w_c = self.T_w_kf @ T_kf_n                               # system.cpp:192
        status = self._vote(np.arange(B), cur_h, T_w_c)            # system.cpp:199
        S = np.nonzero((status == TRACKER_STATE_NEW_KF) & ~self.just_added)[0]
        n_evals, n_pts = np.asarray(out["n_evals"]).copy(), np.asarray(out["n_pts"]).copy()
        self.total_evals += int(n_evals.sum())
        self.total_point_evals += int((n_evals.astype(np.int64) * n_pts.astype(np.int64)).sum())
        self.just_added[:] = False
        if S.size:
            # the previous frame becomes the keyframe and the alignment is repeated against it (system.cpp:203-239)
            old = self.kf_h[S].copy()
            self.be.make_keyframes_h(self.prev_h[S])
            self.kf_h[S] = self.prev_h[S]
            self.T_w_kf[S] = self.T_w_prev[S]
            self.T_kf_prev[S] = eye[S]                             # Pose::setKfFrame: the previous frame IS the keyframe now
            self._clear_up(S)
            o2 = self.be.track_h(self.T_nm1_n[S, :3, :3], self.T_nm1_n[S, :3, 3], self.kf_h[S], cur_h[S])   # system.cpp:225
            T_kf_n[S, :3, :3] = o2["R"]
            T_kf_n[S, :3, 3] = o2["T"]
            T_w_c[S] = self.T_w_kf[S] @ T_kf_n[S]
            status[S] = self._vote(S, cur_h, T_w_c)
            self.just_added[S] = True
            self.n_keyframes += int(S.size)
            self.n_retracks += int(S.size)
            e2, p2 = np.asarray(o2["n_evals"]), np.asarray(o2["n_pts"])
            self.total_evals += int(e2.sum())
            self.total_point_evals += int((e2.astype(np.int64) * p2.astype(np.int64)).sum())
            n_evals[S] += e2
            # keyframes nobody uses any more (a keyframe that is also the previous frame is released below)
            self.be.destroy_h(old[old != self.prev_h[S]])
        self._add_old(np.arange(B), cur_h, T_w_c)                  # system.cpp:253-254
        self.T_nm1_n = _inv(self.T_w_prev) @ T_w_c                 # system.cpp:266  (T_N-1_W * T_W_N)
        # previous frames that did not become keyframes are done
        self.be.destroy_h(self.prev_h[self.prev_h != self.kf_h])
        self.prev_h = cur_h
        self.T_kf_prev = T_kf_n
        self.T_w_prev = T_w_c
        self.T_w_c = T_w_c
        self.status = status
        self.frame += 1
        self.last = dict(out, n_evals=n_evals)
        if self.keep_history:
            self.history.append((T_w_c.copy(), T_kf_n.copy(), n_evals.copy()))
        return self.last

    def _track(self, cur):
        if self.kf_policy == "vote":
            return self._track_vote(cur)
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
        if self.kf_policy == "vote" and getattr(self, "kf_h", None) is not None:
            self.be.destroy_h(np.unique(np.concatenate([self.kf_h, self.prev_h])))
            self.be.destroy_h(self.past_h[self.past_h != 0])
            self.kf_h = self.prev_h = None


class CudaBackend:
    """The product path: everything through the C ABI (revo_b200/api.py)."""

    def __init__(self, ctx, settings, tracker_settings=None, build_ctx=None, track_max_clusters=None):
        """ctx: context (stream) that tracks and promotes keyframes; build_ctx: optional second context whose stream
        uploads frames and builds pyramids, so that the H2D copy of frame k+1 overlaps the tracking of frame k.
        track_max_clusters: resident-cluster cap of the tracking kernel while a second context builds (default
        PIPELINED_TRACK_CLUSTERS; 0 = no cap): at full residency the tracker owns every register of the SMs and the build
        kernels can only run in its tail; with a sixth of the cluster slots left free both run side by side (+2.5 % frames/s on
        B200, profiles/r2_pipeline_cluster_cap.txt).  Without a second context the tracker runs uncapped."""
        from . import api

        self.api = api
        self.ctx = ctx
        self.build_ctx = build_ctx or ctx
        self.settings = settings
        self.tracker = api.TrackerNew(ctx, tracker_settings or api.TrackerSettings(), settings)
        self.campyr = api.CameraPyr(settings)
        cap = PIPELINED_TRACK_CLUSTERS if track_max_clusters is None else track_max_clusters
        ctx.set_track_max_clusters(cap if self.build_ctx is not ctx else 0)

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

    # ---- handle-array flavour (vote policy) ----
    def take(self, batch) -> np.ndarray:
        h = np.frombuffer(batch.arr, dtype=np.uint64).copy()
        batch._alive = False              # the frames are released one by one from now on
        return h

    @staticmethod
    def _arr(h):
        import ctypes as C
        h = np.ascontiguousarray(h, np.uint64)
        return h, h.ctypes.data_as(C.POINTER(C.c_void_p))

    def make_keyframes_h(self, h):
        h, p = self._arr(h)
        if h.size:
            self.ctx.check(self.ctx.lib.revo_pyr_make_keyframe_batch(self.ctx.h, int(h.size), p))

    def destroy_h(self, h):
        h, p = self._arr(h)
        if h.size:
            self.ctx.lib.revo_pyr_destroy_batch(self.ctx.h, int(h.size), p)

    def track_h(self, Rs, Ts, refs_h, curs_h):
        return self.track(Rs, Ts, np.ascontiguousarray(refs_h, np.uint64), np.ascontiguousarray(curs_h, np.uint64))

    def copy_points_h(self, h, lvl):
        h, p = self._arr(h)
        out = np.zeros(h.size, np.uint64)
        if h.size:
            import ctypes as C
            self.ctx.check(self.ctx.lib.revo_pyr_copy_points_batch(self.ctx.h, int(h.size), p, lvl, out.ctypes.data_as(C.POINTER(C.c_void_p))))
        return out

    def vote_h(self, curs_h, n_past, past_h, past_poses, est):
        import ctypes as C
        n = int(len(curs_h))
        status = np.zeros(n, np.int64)
        if n == 0:
            return status
        curs_h, pc = self._arr(curs_h)
        past_h, pp = self._arr(past_h)
        npast = np.ascontiguousarray(n_past, np.int32)
        poses = np.ascontiguousarray(np.asarray(past_poses, np.float32).transpose(0, 1, 3, 2))     # column-major 4x4
        e = np.ascontiguousarray(np.asarray(est, np.float32).transpose(0, 2, 1))
        res = np.zeros(n, self.api.QUALITY_RESULT_DTYPE)
        self.ctx.check(self.ctx.lib.revo_track_quality_batch(self.ctx.h, n, pc, HISTOGRAM_LEVEL, npast.ctypes.data, pp, poses.ctypes.data,
                                                             e.ctypes.data, N_VOTING, res.ctypes.data_as(C.POINTER(self.api.revo_quality_result))))
        self.last_votes = res
        return res["status"].astype(np.int64)
