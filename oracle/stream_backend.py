"""TEST INFRASTRUCTURE (see oracle/revo_oracle.c): the backend object revo_b200.stream.StreamTracker drives, implemented over
the CPU oracle.  Used by bench.py's reference arm / cpu_baseline leg and by the CPU tests; never by the product path."""
import os

import numpy as np


class OracleBackend:
    """The reference's CPU path restated (oracle/): OpenCV kernels through cv2 (all cores), the hand-written
    loops and the tracker through the C port (tracker: OpenMP over independent pairs)."""

    def __init__(self, cam, n_levels):
        from oracle import oracle as O

        import cv2

        self.O = O
        self.orc = O.Oracle("f32")
        # all host threads, explicitly: torch.distributed.run exports OMP_NUM_THREADS=1, which would throttle this arm ~3x
        self.cores = os.cpu_count() or 1
        self.omp_threads = self.orc.set_num_threads(self.cores)
        cv2.setNumThreads(self.cores)
        self.cv2_threads = cv2.getNumThreads()
        self.cfg = O.PyrCfg(n_levels=n_levels)
        self.cam = cam
        self.n_levels = n_levels
        self.ocfg = self.orc.default_cfg()

    def create(self, bgr, depth, n):
        if depth.dtype == np.uint16:     # the reference's reader: depth.convertTo(CV_32FC1, 1.0f / DEPTH_SCALE_FACTOR), iowrapperRGBD.cpp:327
            depth = depth.astype(np.float32) * (np.float32(1.0) / np.float32(5000.0))
        return [self.O.build_pyramid(self.orc, self.cfg, self.cam, bgr[i], depth[i]) for i in range(n)]

    def wait_created(self):
        pass

    def make_keyframes(self, handles):
        for p in handles:
            if not p.dt:
                self.O.make_keyframe(self.orc, p)

    def track(self, Rs, Ts, refs, curs):
        r = self.orc.track_frames_batch(refs, curs, list(Rs), list(Ts), self.ocfg, self.n_levels - 1, 0, True)
        n_pts = np.zeros((len(refs), 6), np.int64)
        for i, c in enumerate(curs):
            for l in range(self.n_levels):
                n_pts[i, l] = len(c.edges3d[l])
        return dict(R=r["R"].astype(np.float32), T=r["T"].astype(np.float32), status=r["status"], n_evals=r["evals"], n_pts=n_pts)

    def destroy(self, handles):
        pass

    # ---- handle-array flavour (StreamTracker kf_policy="vote"): handles are ids into a registry ----
    def take(self, handles) -> np.ndarray:
        if not hasattr(self, "_reg"):
            self._reg, self._next = {}, 1
        ids = np.zeros(len(handles), np.uint64)
        for i, p in enumerate(handles):
            self._reg[self._next] = p
            ids[i] = self._next
            self._next += 1
        return ids

    def _objs(self, h):
        return [self._reg[int(x)] for x in np.asarray(h).reshape(-1)]

    def make_keyframes_h(self, h):
        self.make_keyframes(self._objs(h))

    def destroy_h(self, h):
        for x in np.asarray(h).reshape(-1):
            self._reg.pop(int(x), None)

    def track_h(self, Rs, Ts, refs_h, curs_h):
        return self.track(Rs, Ts, self._objs(refs_h), self._objs(curs_h))

    def copy_points_h(self, h, lvl):
        return self.take([np.array(p.edges3d[lvl], np.float32, copy=True) for p in self._objs(h)])

    def vote_h(self, curs_h, n_past, past_h, past_poses, est):
        status = np.zeros(len(curs_h), np.int64)
        self.last_votes = []
        for i, cur in enumerate(self._objs(curs_h)):
            nv = int(min(n_past[i], 3))
            c = cur.cams[2]
            r = self.O.assess_tracking_quality([self._reg[int(x)] for x in past_h[i, :nv]], list(past_poses[i, :nv]), est[i],
                                               (c.fx, c.fy, c.cx, c.cy, c.w, c.h), cur.depth[2], cur.edges_orig[2])
            self.last_votes.append(r)
            status[i] = r["status"] if nv > 0 else 0
        return status
