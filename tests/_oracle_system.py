"""Oracle-backed stand-ins with the interface revo_b200/system.py drives (test infrastructure only): the reference's
ImgPyramidRGBD / TrackerNew behaviour restated on the CPU through oracle/ (cv2 + C port + numpy vote)."""
import numpy as np

from oracle import oracle as O


class OraclePyr:
    def __init__(self, orc, cfg, cam, bgr, depth, timestamp=0.0):
        self.orc, self.cam = orc, cam
        self.p = O.build_pyramid(orc, cfg, cam, bgr, depth)
        self.timestamp = float(timestamp)
        self.frameId = 0
        self._T = np.eye(4, dtype=np.float32)

    def makeKeyframe(self):
        if not self.p.dt:
            O.make_keyframe(self.orc, self.p)

    def setTwf(self, T):
        self._T = np.asarray(T, np.float32).reshape(4, 4).copy()

    def getTransKFtoWorld(self):
        return self._T

    def returnTimestamp(self):
        return self.timestamp


class OracleTracker:
    histogramLevel = 2

    def __init__(self, orc, n_levels, n_frames_voting=3):
        self.orc, self.n_levels, self.n_frames_voting = orc, n_levels, n_frames_voting
        self.cfg = orc.default_cfg()
        self.past = []
        self.votes = []

    def trackFrames(self, R, T, ref, cur):
        r = self.orc.track_frames(ref.p, cur.p, np.asarray(R, np.float64), np.asarray(T, np.float64), self.cfg, self.n_levels - 1, 0, True)
        return int(r["status"]), r["R"].astype(np.float32), r["T"].astype(np.float32), r["error"]

    def addOldPclAndPose(self, pyr, worldPose, ts=0.0):
        self.past.append((pyr, np.asarray(worldPose, np.float32).copy(), ts))

    def clearUpPastLists(self):
        while len(self.past) > self.n_frames_voting:
            self.past.pop(0)

    def assessTrackingQuality(self, estimatedPose, cur):
        if not self.past:
            return 0
        l = self.histogramLevel
        c = cur.p.cams[l]
        cam = (c.fx, c.fy, c.cx, c.cy, c.w, c.h)
        r = O.assess_tracking_quality([p[0].p.edges3d[l] for p in self.past], [p[1] for p in self.past], estimatedPose, cam,
                                      cur.p.depth[l], cur.p.edges_orig[l], n_frames_voting=self.n_frames_voting)
        self.votes.append(r)
        return r["status"]
