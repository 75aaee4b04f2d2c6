"""Host logic of the REVO::start mirror (revo_b200/system.py: motion model, quality vote, "previous frame becomes keyframe
and track again", pose graph) on the CPU with oracle-backed pyramids / tracker, plus Pose algebra (system.h:89-150)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_pose_algebra():
    from revo_b200.system import Pose, transformFromRT

    class KF:
        def __init__(self, T):
            self.T = T

        def getTransKFtoWorld(self):
            return self.T

    rng = np.random.default_rng(0)
    from revo_b200 import synth

    T_w_kf = synth.se3_exp(rng.normal(0, 0.1, 6)).astype(np.float32)
    T_kf_n = synth.se3_exp(rng.normal(0, 0.1, 6)).astype(np.float32)
    p = Pose(T_kf_n, 1.0, KF(T_w_kf))
    assert np.allclose(p.getCurrToWorld(), T_w_kf @ T_kf_n, atol=1e-6)
    assert np.allclose(p.T_N_W() @ p.T_W_N(), np.eye(4), atol=1e-5)
    p.setKfFrame(KF(p.getCurrToWorld()))                       # the frame becomes its own keyframe
    assert np.array_equal(p.T_kf_N(), np.eye(4, dtype=np.float32))
    assert np.allclose(p.getCurrToWorld(), T_w_kf @ T_kf_n, atol=1e-6)
    M = transformFromRT(np.eye(3), [1, 2, 3])
    assert M[3, 3] == 1 and list(M[:3, 3]) == [1, 2, 3]


def test_revo_loop_follows_ground_truth_and_switches_keyframes():
    from _oracle_system import OraclePyr, OracleTracker
    from oracle import oracle as O
    from revo_b200 import synth
    from revo_b200.system import REVO, TRACKER_STATE_NEW_KF

    w, h, n = 320, 240, 9
    s = synth.make_stream(77, n, w, h, max_trans=0.03, max_rot_deg=1.5)
    cam = synth.intrinsics(w, h)
    orc = O.Oracle("f32")
    cfg = O.PyrCfg(n_levels=3)
    trk = OracleTracker(orc, 3)
    sysm = REVO(trk)
    for i in range(n):
        sysm.processFrame(OraclePyr(orc, cfg, cam, *s["frames"][i], timestamp=0.033 * i))
    traj = sysm.trajectory()
    assert traj.shape == (n, 4, 4) and sysm.noFrames == n and len(sysm.mPoseGraph) == n
    for i in (n // 2, n - 1):
        T_gt = np.linalg.inv(s["T_w_c"][0]) @ s["T_w_c"][i]
        D = np.linalg.inv(T_gt) @ traj[i].astype(np.float64)
        assert synth.rot_angle(D[:3, :3]) < 8e-3 and np.linalg.norm(D[:3, 3]) < 2e-2, (i, D)
    # the vote ran on every frame after the first; whenever it asked for a keyframe (and the previous frame was not one
    # already) the previous frame was promoted and the frame tracked again against it
    assert len(trk.votes) >= n - 1
    assert sysm.nKeyFrames == 1 + len(sysm.retracked)
    for fid in sysm.retracked:
        assert sysm.mPoseGraph[fid].kfFrame is sysm.mPoseGraph[fid - 1].kfFrame           # parent = the promoted previous frame
        assert np.array_equal(sysm.mPoseGraph[fid - 1].T_kf_N(), np.eye(4, dtype=np.float32))
    # motion model: the initial guess of the next frame is T_kf_N * T_NM1_N (system.cpp:268)
    T_init = sysm.mPoseGraph[-1].T_kf_N() @ sysm.T_NM1_N
    assert np.allclose(T_init[:3, :3], sysm.R, atol=1e-6) and np.allclose(T_init[:3, 3], sysm.T, atol=1e-6)


def test_forced_keyframe_switch():
    """A tracker whose vote always asks for a keyframe: promotion happens on every second frame only
    (justAddedNewKeyframe, system.cpp:203,240-243)."""
    from revo_b200.system import REVO, TRACKER_STATE_NEW_KF

    class P:
        def __init__(self, ts):
            self.ts, self.kf, self.T, self.frameId = ts, False, np.eye(4, dtype=np.float32), 0

        def makeKeyframe(self):
            self.kf = True

        def setTwf(self, T):
            self.T = np.asarray(T, np.float32)

        def getTransKFtoWorld(self):
            return self.T

        def returnTimestamp(self):
            return self.ts

    class Trk:
        histogramLevel = 2
        cleared = 0

        def trackFrames(self, R, T, ref, cur):
            assert ref.kf
            return 0, np.eye(3, dtype=np.float32), np.array([0.01, 0, 0], np.float32), 0.1

        def assessTrackingQuality(self, T, cur):
            return TRACKER_STATE_NEW_KF

        def addOldPclAndPose(self, *a):
            pass

        def clearUpPastLists(self):
            Trk.cleared += 1

    sysm = REVO(Trk())
    frames = [P(i) for i in range(6)]
    for f in frames:
        sysm.processFrame(f)
    # frame 0 is a keyframe; frame 1: vote NEW_KF but justAdded -> no switch; frame 2: switch to frame 1; frame 3: no; frame 4: switch to 3
    assert [f.kf for f in frames] == [True, True, False, True, False, False]
    assert sysm.retracked == [2, 4] and Trk.cleared == 2 and sysm.nKeyFrames == 3
    # the fake tracker always reports 1 cm relative to the keyframe in use: world x = keyframe x + 1 cm
    assert np.allclose(sysm.trajectory()[:, 0, 3], [0.0, 0.01, 0.02, 0.02, 0.03, 0.03], atol=1e-6)


def test_multi_stream_driver_equals_independent_runs():
    """MultiStreamREVO batches the trackFrames requests of B streams (first alignment of all streams in one call, the
    re-tracks after a keyframe switch in a second one) and must reproduce B separate REVO runs exactly."""
    from _oracle_system import OraclePyr, OracleTracker
    from oracle import oracle as O
    from revo_b200 import synth
    from revo_b200.system import REVO, MultiStreamREVO

    w, h, n, B = 160, 120, 7, 3
    cam = synth.intrinsics(w, h)
    orc = O.Oracle("f32")
    cfg = O.PyrCfg(n_levels=3)
    # stream 1 moves fast, so that its vote asks for keyframes while the others keep theirs
    streams = [synth.make_stream(300 + b, n, w, h, max_trans=0.01 + 0.04 * (b == 1), max_rot_deg=0.5 + 2.5 * (b == 1)) for b in range(B)]

    def pyr(b, i):
        return OraclePyr(orc, cfg, cam, *streams[b]["frames"][i], timestamp=0.033 * i)

    single = [REVO(OracleTracker(orc, 3)) for _ in range(B)]
    for b in range(B):
        for i in range(n):
            single[b].processFrame(pyr(b, i))

    trackers = [OracleTracker(orc, 3) for _ in range(B)]
    calls = []

    def track_batch(requests):
        calls.append(len(requests))
        return [trackers[0].trackFrames(*r) for r in requests]       # trackFrames itself is stateless

    multi = MultiStreamREVO(trackers, track_batch)
    for i in range(n):
        poses = multi.processFrames([pyr(b, i) for b in range(B)])
        assert len(poses) == B and all(p.shape == (4, 4) for p in poses)
    for b in range(B):
        assert np.array_equal(multi.trajectories()[b], single[b].trajectory())
        assert multi.streams[b].retracked == single[b].retracked
        assert multi.streams[b].nKeyFrames == single[b].nKeyFrames
    # the first frame needs no tracking; afterwards one call with all B streams, plus one per frame for the re-tracks
    n_retracks = sum(len(s.retracked) for s in single)
    assert n_retracks >= 1                                          # the second, smaller batch is exercised
    assert calls.count(B) >= n - 1 and sum(calls) == B * (n - 1) + n_retracks
    assert multi.batch_sizes == calls


def test_cuda_track_batch_adapter_conventions():
    """cuda_track_batch: requests -> one trackFramesBatch call; the column-major R of the result record comes back as a
    row-major 3x3 (what REVO._frame composes poses with)."""
    from revo_b200 import api
    from revo_b200.system import cuda_track_batch

    R_true = np.arange(9, dtype=np.float32).reshape(3, 3)

    class FakeTracker:
        def trackFramesBatch(self, Rs, Ts, refs, curs):
            assert Rs.shape == (2, 3, 3) and Ts.shape == (2, 3) and refs == ["k0", "k1"] and curs == ["c0", "c1"]
            out = np.zeros(2, api.TRACK_RESULT_DTYPE)
            for i in range(2):
                out["R"][i] = api._R_to_c(R_true + i)            # column-major, as the C ABI returns it
                out["t"][i] = Ts[i] + 1
                out["status"][i] = 2 * i
                out["error"][i] = 0.5 + i
            return out

    res = cuda_track_batch(FakeTracker())([(np.eye(3), np.zeros(3), "k0", "c0"), (np.eye(3), np.ones(3), "k1", "c1")])
    assert [r[0] for r in res] == [0, 2] and [r[3] for r in res] == [0.5, 1.5]
    assert np.array_equal(res[0][1], R_true) and np.array_equal(res[1][1], R_true + 1)
    assert np.array_equal(res[1][2], [2, 2, 2])


def test_stream_tracker_vote_policy_equals_the_per_stream_main_loop():
    """StreamTracker(kf_policy="vote") -- the reference's keyframe policy (system.cpp:199-239) vectorised over the streams with
    handle arrays, batched votes and batched point-list copies -- against B separate REVO main loops over the same oracle
    objects: same keyframe decisions, same re-tracks, same world poses."""
    from _oracle_system import OraclePyr, OracleTracker
    from oracle import oracle as O
    from oracle.stream_backend import OracleBackend
    from revo_b200 import synth
    from revo_b200.stream import StreamTracker
    from revo_b200.system import REVO

    w, h, n, B = 160, 120, 9, 3
    cam = synth.intrinsics(w, h)
    streams = [synth.make_stream(300 + b, n, w, h, max_trans=0.01 + 0.04 * (b == 1), max_rot_deg=0.5 + 2.5 * (b == 1)) for b in range(B)]
    orc = O.Oracle("f32")
    cfg = O.PyrCfg(n_levels=3)
    single = [REVO(OracleTracker(orc, 3)) for _ in range(B)]
    for b in range(B):
        for i in range(n):
            single[b].processFrame(OraclePyr(orc, cfg, cam, *streams[b]["frames"][i], timestamp=0.033 * i))

    be = OracleBackend(cam, 3)
    st = StreamTracker(be, B, kf_policy="vote")
    st.keep_history = True
    frames = lambda i: (np.stack([streams[b]["frames"][i][0] for b in range(B)]), np.stack([streams[b]["frames"][i][1] for b in range(B)]))
    st.start(*frames(0))
    kf_flags = []
    for i in range(1, n):
        st.step(*frames(i))
        kf_flags.append(st.just_added.copy())
    assert st.n_retracks == sum(len(s.retracked) for s in single) >= 1
    for b in range(B):
        assert [i + 1 for i, f in enumerate(kf_flags) if f[b]] == single[b].retracked
        traj = single[b].trajectory()
        for i in range(1, n):
            assert np.array_equal(st.history[i - 1][0][b], traj[i]), (b, i)
    # the vote history stays bounded (first three entries vote, last three survive a clear-up)
    assert st.n_past.max() <= 6 and len(be._reg) <= B * (2 + 6)
    st.close()
    assert not be._reg
