"""Host-side mirror of the per-frame body of ``REVO::start`` (system/system.cpp:128-283) and of ``REVO::Pose``
(system/system.h:89-150): motion-model initialisation, the tracking-quality vote and the "take the previous frame as
keyframe and track again" policy, pose-graph bookkeeping.  It is the caller of the hot path, not part of it: pure host
logic over objects with the reference's interface

    pyramid:  makeKeyframe(), setTwf(T), getTransKFtoWorld(), returnTimestamp(), frameId
    tracker:  trackFrames(R, T, ref, cur) -> (status, R, T, error), assessTrackingQuality(T_w_c, cur) -> status,
              addOldPclAndPose(pyr, T_w_c, ts), clearUpPastLists(), histogramLevel

so the same code drives the CUDA classes of :mod:`revo_b200.api` and, in the CPU tests, oracle-backed stand-ins.
IO, viewer, logging and pose output of the reference loop are out of scope.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

TRACKER_STATE_OK, TRACKER_STATE_LOST, TRACKER_STATE_NEW_KF, TRACKER_STATE_UNKNOWN = range(4)


class VoteRequest(tuple):
    """``(estimatedPose, currPyr)``: an ``assessTrackingQuality`` request of the per-frame coroutine (a ``trackFrames`` request is
    a plain 4-tuple ``(R, T, refPyr, currPyr)``)."""
    __slots__ = ()


def transformFromRT(R, T) -> np.ndarray:
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = np.asarray(R, np.float32).reshape(3, 3)
    M[:3, 3] = np.asarray(T, np.float32).reshape(3)
    return M


def _inv(T: np.ndarray) -> np.ndarray:
    return np.linalg.inv(T.astype(np.float64)).astype(np.float32)


class Pose:
    """``REVO::Pose`` (system/system.h:89-150): pose of a frame relative to its parent keyframe."""

    def __init__(self, T_kf_curr, timestamp: float, kfFrame):
        self.T_kf_curr = np.asarray(T_kf_curr, np.float32).reshape(4, 4).copy()
        self.timestamp = float(timestamp)
        self.kfFrame = kfFrame

    def getCurrToWorld(self) -> np.ndarray:          # T_W_curr = T_W_KF * T_KF_CURR  (system.h:131-134)
        return (self.kfFrame.getTransKFtoWorld().astype(np.float32) @ self.T_kf_curr).astype(np.float32)

    def T_W_N(self) -> np.ndarray:
        return self.getCurrToWorld()

    def T_N_W(self) -> np.ndarray:
        return _inv(self.getCurrToWorld())

    def T_kf_N(self) -> np.ndarray:
        return self.T_kf_curr

    def setKfFrame(self, kfFrame):                   # only called when the "previous frame" becomes keyframe (system.h:140-146)
        self.kfFrame = kfFrame
        self.T_kf_curr = np.eye(4, dtype=np.float32)

    def returnTimestamp(self) -> float:
        return self.timestamp


class REVO:
    """The tracking part of ``REVO::start`` for one stream: feed pyramids in order with :meth:`processFrame`."""

    def __init__(self, tracker):
        self.mTracker = tracker
        self.mPoseGraph: List[Pose] = []
        self.kfPyr = None
        self.prevPyr = None
        self.noFrames = 0
        self.nKeyFrames = 0
        self.justAddedNewKeyframe = False
        self.R = np.eye(3, dtype=np.float32)        # initial guess of the next frame relative to the keyframe
        self.T = np.zeros(3, dtype=np.float32)
        self.T_NM1_N = np.eye(4, dtype=np.float32)
        self.trackerStatus = TRACKER_STATE_UNKNOWN
        self.error = 0.0
        self.retracked: List[int] = []               # frame ids at which the previous frame was promoted and tracking repeated

    def _frame(self, currPyr, pcl=None):
        """One iteration of the ``while`` loop (system.cpp:147-275) as a coroutine: it YIELDS every ``trackFrames`` request
        ``(R_init, T_init, refPyr, currPyr)`` and is sent the result ``(status, R, T, error)``, and every
        ``assessTrackingQuality`` request (:class:`VoteRequest`) and is sent the status, so that a driver can batch the
        requests of many streams into one launch each (:class:`MultiStreamREVO`).  ``pcl``: what to hand to
        ``addOldPclAndPose`` for this frame (default: the pyramid; a driver may have copied the lists of all streams in one
        call).  Returns the frame's world pose."""
        trk = self.mTracker
        pcl = currPyr if pcl is None else pcl
        currPyr.frameId = self.noFrames
        if self.noFrames == 0:                       # first frame -> keyframe (system.cpp:151-175)
            self.kfPyr = self.prevPyr = currPyr
            currPyr.makeKeyframe()
            currPyr.setTwf(np.eye(4, dtype=np.float32))
            self.mPoseGraph.append(Pose(np.eye(4), currPyr.returnTimestamp(), currPyr))
            self.nKeyFrames += 1
            self.noFrames += 1
            self.justAddedNewKeyframe = True
            trk.addOldPclAndPose(pcl, np.eye(4, dtype=np.float32), currPyr.returnTimestamp())
            return np.eye(4, dtype=np.float32)
        self.noFrames += 1
        status, R, T, self.error = yield (self.R, self.T, self.kfPyr, currPyr)                           # :188
        T_KF_N = transformFromRT(R, T)
        currPoseInWorld = (self.kfPyr.getTransKFtoWorld().astype(np.float32) @ T_KF_N).astype(np.float32)   # :192
        status = yield VoteRequest((currPoseInWorld, currPyr))                                           # :199
        if status == TRACKER_STATE_NEW_KF and not self.justAddedNewKeyframe:
            # tracking gets inaccurate: take the previous frame as keyframe and optimise again (:203-239)
            self.kfPyr = self.prevPyr
            self.kfPyr.setTwf(self.mPoseGraph[-1].getCurrToWorld())
            self.kfPyr.makeKeyframe()
            self.mPoseGraph[-1].setKfFrame(self.kfPyr)
            self.nKeyFrames += 1
            trk.clearUpPastLists()
            _, R, T, self.error = yield (self.T_NM1_N[:3, :3], self.T_NM1_N[:3, 3], self.kfPyr, currPyr)   # :225
            T_KF_N = transformFromRT(R, T)
            currPoseInWorld = (self.kfPyr.getTransKFtoWorld().astype(np.float32) @ T_KF_N).astype(np.float32)
            status = yield VoteRequest((currPoseInWorld, currPyr))
            self.justAddedNewKeyframe = True
            self.retracked.append(currPyr.frameId)
        else:
            self.justAddedNewKeyframe = False
        self.trackerStatus = status
        # add the frame to the pose graph, remember its edge cloud for the vote (:253-254)
        self.mPoseGraph.append(Pose(T_KF_N, currPyr.returnTimestamp(), self.kfPyr))
        trk.addOldPclAndPose(pcl, currPoseInWorld, currPyr.returnTimestamp())
        # relative motion N-1 -> N and the constant-velocity guess for the next frame (:262-271)
        self.T_NM1_N = (self.mPoseGraph[-2].T_N_W() @ self.mPoseGraph[-1].T_W_N()).astype(np.float32)
        T_init = (self.mPoseGraph[-1].T_kf_N() @ self.T_NM1_N).astype(np.float32)
        self.R, self.T = T_init[:3, :3].copy(), T_init[:3, 3].copy()
        self.prevPyr = currPyr
        return self.mPoseGraph[-1].getCurrToWorld()

    def processFrame(self, currPyr) -> np.ndarray:
        """One iteration of the ``while`` loop (system.cpp:147-275). Returns the frame's pose in the world."""
        g = self._frame(currPyr)
        try:
            req = next(g)
            while True:
                if isinstance(req, VoteRequest):
                    req = g.send(self.mTracker.assessTrackingQuality(*req))
                else:
                    req = g.send(self.mTracker.trackFrames(*req))
        except StopIteration as done:
            return done.value

    def trajectory(self) -> np.ndarray:
        """(n_frames, 4, 4) world poses of all frames processed so far."""
        return np.stack([p.getCurrToWorld() for p in self.mPoseGraph])


class MultiStreamREVO:
    """B independent streams, each with the per-frame logic of :class:`REVO` (its own pose graph, motion model and vote
    history), advancing one frame per call with the ``trackFrames`` requests of all streams batched: one launch for the
    first alignment of every stream, one more for the streams whose vote asked for a new keyframe.  By construction the
    result equals B separate :class:`REVO` runs (same coroutine), which is what the tests check.

    trackers:    one tracker object per stream (vote state: ``assessTrackingQuality`` / ``addOldPclAndPose`` /
                 ``clearUpPastLists``)
    track_batch: ``f(requests) -> results`` with ``requests = [(R, T, refPyr, currPyr), ...]`` and
                 ``results = [(status, R, T, error), ...]`` in the same order
    vote_batch:  optional ``f(trackers, estimatedPoses, currPyrs) -> statuses`` (``api.assess_tracking_quality_batch``: the votes
                 of all streams in one launch pair); default: every tracker's own ``assessTrackingQuality``
    pcl_batch:   optional ``f(currPyrs) -> objects`` handed to ``addOldPclAndPose`` in place of the pyramids (one batched copy
                 of the level-``histogramLevel`` lists, ``api.copy_point_lists``)
    """

    def __init__(self, trackers, track_batch, vote_batch=None, pcl_batch=None):
        self.streams: List[REVO] = [REVO(t) for t in trackers]
        self.track_batch = track_batch
        self.vote_batch = vote_batch
        self.pcl_batch = pcl_batch
        self.batch_sizes: List[int] = []             # trackFrames requests per launch, for tests / diagnostics
        self.vote_batch_sizes: List[int] = []

    def processFrames(self, currPyrs) -> List[np.ndarray]:
        assert len(currPyrs) == len(self.streams)
        pcls = self.pcl_batch(currPyrs) if self.pcl_batch else [None] * len(currPyrs)
        gens = [s._frame(p, c) for s, p, c in zip(self.streams, currPyrs, pcls)]
        poses: List[Optional[np.ndarray]] = [None] * len(gens)
        pending = {}
        for i, g in enumerate(gens):
            try:
                pending[i] = next(g)
            except StopIteration as done:
                poses[i] = done.value
        while pending:
            tracks = sorted(i for i in pending if not isinstance(pending[i], VoteRequest))
            votes = sorted(i for i in pending if isinstance(pending[i], VoteRequest))
            answers = {}
            if tracks:
                answers.update(zip(tracks, self.track_batch([pending[i] for i in tracks])))
                self.batch_sizes.append(len(tracks))
            if votes:
                trks = [self.streams[i].mTracker for i in votes]
                if self.vote_batch:
                    st = self.vote_batch(trks, [pending[i][0] for i in votes], [pending[i][1] for i in votes])
                else:
                    st = [t.assessTrackingQuality(*pending[i]) for t, i in zip(trks, votes)]
                answers.update(zip(votes, st))
                self.vote_batch_sizes.append(len(votes))
            pending = {}
            for i, r in answers.items():
                try:
                    pending[i] = gens[i].send(r)
                except StopIteration as done:
                    poses[i] = done.value
        return poses

    def trajectories(self) -> List[np.ndarray]:
        return [s.trajectory() for s in self.streams]


def cuda_track_batch(tracker):
    """``track_batch`` for :class:`MultiStreamREVO` over ``revo_b200.api.TrackerNew.trackFramesBatch``."""

    def run(requests):
        n = len(requests)
        Rs = np.stack([np.asarray(r[0], np.float32).reshape(3, 3) for r in requests])
        Ts = np.stack([np.asarray(r[1], np.float32).reshape(3) for r in requests])
        out = tracker.trackFramesBatch(Rs, Ts, [r[2] for r in requests], [r[3] for r in requests])
        return [(int(out["status"][i]), out["R"][i].reshape(3, 3).T.copy(), out["t"][i].copy(), float(out["error"][i]))
                for i in range(n)]

    return run
