"""GPU parity: the fused residual/Jacobian/normal-equation evaluation, Optimizer::trackFrames and
TrackerNew::trackFrames through the C ABI vs the CPU oracle (float32 "reference-as-is" and float64 "truth").

Tolerances (floating point; stated per test):
 * one evaluation record: |gpu - f64| <= 2e-5 * scale per block (A, b, sums), counts exact;
 * pose after the SAME number of LM tries (fixed-iteration mode): <= 1e-4 rad, <= 1e-4 m vs the f64 oracle;
 * default termination rules: GPU pose must agree with the oracle to within the oracle's own f32-vs-f64 spread
   (the accept/convergence tests sit on float-rounding knife edges, SURVEY.md 7 "Hard parts"), and to <= 1e-4
   whenever the LM traces coincide.
"""
import os

import numpy as np
import pytest

from conftest import rot_angle, synth_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def engine(ctx):
    """The library has one tracking engine (one thread-block cluster per pair, track.cu)."""
    return "cluster"


def _settings(cam, n_levels):
    from revo_b200 import api

    fx, fy, cx, cy, w, h = cam
    return api.ImgPyramidSettings(PYR_MIN_LVL=n_levels - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)


def _build(ctx, orc, p, n_levels):
    """Oracle pyramids + GPU pyramids carrying the ORACLE's arrays (so optimizer parity is isolated
    from pyramid parity)."""
    from oracle import oracle as O
    from revo_b200 import api

    cfg = O.PyrCfg(n_levels=n_levels)
    ok = O.build_pyramid(orc, cfg, p["cam"], *p["key"])
    O.make_keyframe(orc, ok)
    oc = O.build_pyramid(orc, cfg, p["cam"], *p["cur"])
    st = _settings(p["cam"], n_levels)
    gk = api.ImgPyramidRGBD(ctx, st, None, *p["key"])
    gc = api.ImgPyramidRGBD(ctx, st, None, *p["cur"])
    for l in range(n_levels):
        gk.uploadLevel(l, pts4=ok.edges3d[l], dt=ok.dt[l], opt4=ok.opt[l])
        gc.uploadLevel(l, pts4=oc.edges3d[l])
    return ok, oc, gk, gc, st


def _rec_close(g, o, tol=2e-5, max_flips=0, unw_tol=None):
    """Record parity.  Counts must agree exactly unless max_flips > 0 (poses that put points exactly on the
    u>1 / v>1 / u<w-2 / v<h-2 bounds, e.g. the identity, classify border pixels on float rounding)."""
    flips = abs(g[29] - o[29])
    assert flips <= max_flips and g[29] + g[30] == o[29] + o[30], (g[29:31], o[29:31])
    if flips:
        tol = max(tol, 3.0 * flips / max(o[29], 1.0))
    sA = np.abs(o[:21]).max() + 1e-30
    sb = np.abs(o[:21]).max() ** 0.5 * np.abs(o[27]) ** 0.5 + 1e-30   # |sum w r v| <= sqrt(sum w v^2 * sum w r^2)
    assert np.abs(g[:21] - o[:21]).max() <= tol * sA, (g[:21], o[:21])
    assert np.abs(g[21:27] - o[21:27]).max() <= tol * sb, (g[21:27], o[21:27])
    assert abs(g[27] - o[27]) <= tol * abs(o[27]) + 1e-12
    assert abs(g[28] - o[28]) <= (unw_tol or tol) * abs(o[28]) + 1e-12


@pytest.mark.parametrize("seed", [1, 21])
def test_eval_record_matches_oracle(ctx, orc32, orc64, seed):
    from revo_b200 import api, synth

    p = synth_pair(seed)
    ok, oc, gk, gc, st = _build(ctx, orc64, p, 3)
    opt = api.Optimizer(ctx, api.OptimizerSettings(USE_EDGE_FILTER=True))
    ocfg = orc64.default_cfg()
    near = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
    far = synth.se3_exp([0.05, -0.03, 0.02, 0.02, -0.03, 0.01])
    poses = [(near[:3, :3], near[:3, 3]), (p["T_kf_cur"][:3, :3], p["T_kf_cur"][:3, 3]), (far[:3, :3], far[:3, 3])]
    for lvl in range(3):
        for R, T in poses:
            R32, T32 = np.asarray(R, np.float32), np.asarray(T, np.float32)
            g = opt.evalRecord(gk, gc, R32, T32, lvl)
            o = orc64.eval_record(oc.edges3d[lvl], ok.opt[lvl], oc.cams[lvl], R32, T32, ocfg, lvl)
            _rec_close(g, o, max_flips=1)
            # and the float32 sequential reference-as-is agrees with both at its own precision
            o32 = orc32.eval_record(oc.edges3d[lvl], ok.opt[lvl], oc.cams[lvl], R32, T32, ocfg, lvl)
            _rec_close(o32, o, tol=2e-3, max_flips=1, unw_tol=0.06)
    # identity pose: every point projects onto an integer pixel, so border pixels sit exactly on the bounds
    # test and classify on rounding (the f32 and f64 oracles disagree with each other there, too)
    I, Z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    g = opt.evalRecord(gk, gc, I, Z, 0)
    o = orc64.eval_record(oc.edges3d[0], ok.opt[0], oc.cams[0], I, Z, ocfg, 0)
    o32 = orc32.eval_record(oc.edges3d[0], ok.opt[0], oc.cams[0], I, Z, ocfg, 0)
    _rec_close(g, o, max_flips=200)
    _rec_close(o32, o, tol=2e-3, max_flips=200, unw_tol=0.06)
    # edge filter off
    opt2 = api.Optimizer(ctx, api.OptimizerSettings(USE_EDGE_FILTER=False))
    ocfg.use_edge_filter = 0
    R32, T32 = np.asarray(near[:3, :3], np.float32), np.asarray(near[:3, 3], np.float32)
    g = opt2.evalRecord(gk, gc, R32, T32, 0)
    o = orc64.eval_record(oc.edges3d[0], ok.opt[0], oc.cams[0], R32, T32, ocfg, 0)
    _rec_close(g, o, max_flips=1)


def test_eval_is_deterministic_and_shape_independent(ctx, orc64):
    """Same record bit-for-bit run to run; different cluster shapes only differ by summation order."""
    from revo_b200 import api

    p = synth_pair(1)
    ok, oc, gk, gc, st = _build(ctx, orc64, p, 3)
    opt = api.Optimizer(ctx, api.OptimizerSettings(USE_EDGE_FILTER=True))
    R, T = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    base = None
    try:
        for ctas, thr in [(0, 0), (1, 256), (2, 512), (4, 128), (8, 1024), (16, 512)]:
            ctx.set_track_shape(ctas, thr)
            a = opt.evalRecord(gk, gc, R, T, 0)
            b = opt.evalRecord(gk, gc, R, T, 0)
            assert np.array_equal(a, b), (ctas, thr)
            if base is None:
                base = a
            _rec_close(a, base, tol=1e-5)
    finally:
        ctx.set_track_shape(0, 0)


@pytest.mark.parametrize("seed,n_tries,n_levels,w,h", [(1, 8, 3, 640, 480), (22, 5, 3, 640, 480), (31, 6, 4, 640, 480),
                                                        (3, 5, 5, 1280, 960)])
def test_track_level_fixed_iterations(ctx, orc64, seed, n_tries, n_levels, w, h):
    """Same iteration count on both sides (north_star: "after the same iteration count"): <= 1e-4 rad / 1e-4 m, at 3, 4
    (BASELINE configs[1]) and 5 levels (configs[2], 1280x960)."""
    from revo_b200 import api, synth

    p = synth_pair(seed, w, h)
    ok, oc, gk, gc, st = _build(ctx, orc64, p, n_levels)
    # start off the integer pixel grid (see test_eval_record_matches_oracle) so both sides classify identically
    T0 = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
    R, T = np.asarray(T0[:3, :3], np.float32), np.asarray(T0[:3, 3], np.float32)
    Ro, To = R.copy(), T.copy()
    for lvl in range(n_levels - 1, -1, -1):
        s = api.OptimizerSettings(USE_EDGE_FILTER=True, max_lm_tries=n_tries, convergenceEps=[2.0] * 6)
        opt = api.Optimizer(ctx, s)
        ri = api.ResidualInfo()
        err, R, T = opt.trackFrames(gk, gc, R, T, lvl, ri)
        ocfg = orc64.default_cfg()
        for l in range(6):
            ocfg.convergence_eps[l] = 2.0
        r = orc64.track_level(oc.edges3d[lvl], ok.opt[lvl], oc.cams[lvl], Ro, To, ocfg, lvl, max_tries=n_tries)
        Ro, To = r["R"].astype(np.float32), r["T"].astype(np.float32)
        assert opt.last_n_evals == r["n_evals"], (lvl, opt.last_n_evals, r["n_evals"])
        assert rot_angle(R, Ro) <= 1e-4, (lvl, rot_angle(R, Ro))
        assert np.linalg.norm(T - To) <= 1e-4, (lvl, np.linalg.norm(T - To))
        assert abs(err - r["error"]) <= 1e-4 * max(1.0, abs(r["error"]))
        assert ri.goodPtsEdges == r["good"] and ri.badPtsEdges == r["bad"]


@pytest.mark.parametrize("seed", [1, 23, 24])
def test_track_frames_default_rules(ctx, orc32, orc64, seed):
    """TrackerNew::trackFrames with the reference's own termination rules."""
    from revo_b200 import api, synth

    xi = synth.XI_CONFIG1 if seed == 1 else None
    p = synth_pair(seed, xi=xi)
    ok, oc, gk, gc, st = _build(ctx, orc64, p, 3)
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    out, traces = trk.trackFramesBatch([np.eye(3)], [np.zeros(3)], [gk], [gc], trace_cap=256)
    Rg, Tg = api.result_R(out[0]), out[0]["t"]
    r32 = orc32.track_frames(ok, oc, np.eye(3), np.zeros(3), orc32.default_cfg(), 2, 0, True)
    r64 = orc64.track_frames(ok, oc, np.eye(3), np.zeros(3), orc64.default_cfg(), 2, 0, True)
    spread_r = rot_angle(r32["R"], r64["R"])
    spread_t = np.linalg.norm(r32["T"].astype(np.float64) - r64["T"])
    d_r = min(rot_angle(Rg, r32["R"]), rot_angle(Rg, r64["R"]))
    d_t = min(np.linalg.norm(Tg - r32["T"]), np.linalg.norm(Tg - r64["T"]))
    # converged to the ground truth as well as the oracle does
    Tgt = p["T_kf_cur"]
    gt_r, gt_t = rot_angle(Rg, Tgt[:3, :3]), np.linalg.norm(Tg - Tgt[:3, 3])
    o_r, o_t = rot_angle(r64["R"], Tgt[:3, :3]), np.linalg.norm(r64["T"] - Tgt[:3, 3])
    print(f"seed {seed}: gpu evals {list(out[0]['n_evals'][:3])} f32 {r32['evals'][:3]} f64 {r64['evals'][:3]} "
          f"d_r {d_r:.2e} d_t {d_t:.2e} spread {spread_r:.2e}/{spread_t:.2e} gt {gt_r:.2e}/{gt_t:.2e} oracle-gt {o_r:.2e}/{o_t:.2e}")
    assert out[0]["rc"] == 0 and out[0]["status"] == r64["status"]
    assert d_r <= max(1e-4, 1.5 * spread_r) and d_t <= max(1e-4, 1.5 * spread_t)
    assert gt_r <= o_r + 3e-4 and gt_t <= o_t + 1e-3
    same_trace = list(out[0]["n_evals"][:3]) == r64["evals"][:3]
    if same_trace:
        assert rot_angle(Rg, r64["R"]) <= 1e-4 and np.linalg.norm(Tg - r64["T"]) <= 1e-4


def test_track_batch_matches_single_and_check_init(ctx, orc64):
    """Batch of pairs == the same pairs one by one (bitwise); a bad initial pose is reset by checkInitializationValues."""
    from revo_b200 import api, synth

    ps = [synth_pair(s) for s in (1, 23)]
    built = [_build(ctx, orc64, p, 3) for p in ps]
    st = built[0][4]
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    bad = synth.se3_exp([0.4, 0.3, -0.2, 0.2, -0.15, 0.1])
    Rs = [np.eye(3), bad[:3, :3]]
    Ts = [np.zeros(3), bad[:3, 3]]
    out = trk.trackFramesBatch(Rs, Ts, [b[2] for b in built], [b[3] for b in built])
    for i in range(2):
        st_i, R_i, T_i, e_i = trk.trackFrames(Rs[i], Ts[i], built[i][2], built[i][3])
        assert np.array_equal(api.result_R(out[i]), R_i) and np.array_equal(out[i]["t"], T_i)
        assert out[i]["status"] == st_i
    assert out[0]["used_identity_init"] == 0 and out[1]["used_identity_init"] == 1
    ok, oc = built[1][0], built[1][1]
    r = orc64.track_frames(ok, oc, bad[:3, :3], bad[:3, 3], orc64.default_cfg(), 2, 0, True)
    assert rot_angle(api.result_R(out[1]), r["R"]) <= 5e-4


def test_track_error_codes(ctx, orc64):
    from revo_b200 import api

    p = synth_pair(1)
    ok, oc, gk, gc, st = _build(ctx, orc64, p, 3)
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    # non-orthogonal R: the reference abort()s inside Sophus; here an error code
    with pytest.raises(api.RevoError) as ei:
        trk.trackFrames(np.eye(3) * 1.1, np.zeros(3), gk, gc)
    assert ei.value.code == 5
    # reference frame that is not a keyframe
    with pytest.raises(api.RevoError) as ei:
        trk.trackFrames(np.eye(3), np.zeros(3), gc, gk)
    assert ei.value.code == 4


def test_end_to_end_gpu_pyramids_track_to_ground_truth(ctx, orc64):
    """Pyramids built by the CUDA path (not uploaded): full pipeline converges like the oracle."""
    from revo_b200 import api, synth

    p = synth_pair(1, xi=synth.XI_CONFIG1)
    st = _settings(p["cam"], 3)
    gk = api.ImgPyramidRGBD(ctx, st, None, *p["key"])
    gc = api.ImgPyramidRGBD(ctx, st, None, *p["cur"])
    gk.makeKeyframe()
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    status, R, T, err = trk.trackFrames(np.eye(3), np.zeros(3), gk, gc)
    Tgt = p["T_kf_cur"]
    assert status == api.TRACKER_STATE_OK
    assert rot_angle(R, Tgt[:3, :3]) < 1.5e-3 and np.linalg.norm(T - Tgt[:3, 3]) < 3e-3
    assert sum(trk.last_result.n_evals) >= 6


def test_tracking_quality_vote(ctx, orc64):
    """TrackerNew::assessTrackingQuality (tracker.cpp:118-201) on the device vs the numpy restatement: exact counts."""
    from oracle import oracle as O
    from revo_b200 import api, synth

    seeds = (1, 23, 24)
    ps = [synth_pair(s) for s in seeds]
    st = _settings(ps[0]["cam"], 3)
    pyrs = [api.ImgPyramidRGBD(ctx, st, None, *p["key"]) for p in ps]
    cur = api.ImgPyramidRGBD(ctx, st, None, *ps[0]["cur"])
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    rng = np.random.default_rng(3)
    poses = [synth.se3_exp(rng.normal(0, 0.01, 6)) for _ in seeds]
    est = synth.se3_exp(rng.normal(0, 0.01, 6))
    assert trk.assessTrackingQuality(est, cur) == api.TRACKER_STATE_OK and trk.last_quality is None     # nothing to vote with
    lvl = trk.histogramLevel
    cam2 = cur._cam(lvl)
    cam = (cam2.fx, cam2.fy, cam2.cx, cam2.cy, cam2.width, cam2.height)
    depth, edges = cur.returnDepth(lvl), cur.returnOrigEdges(lvl)
    for k, (p, T) in enumerate(zip(pyrs, poses)):
        trk.addOldPclAndPose(p, T, float(k))
        status = trk.assessTrackingQuality(est, cur)
        q = trk.last_quality
        o = O.assess_tracking_quality([x.return3DEdges(lvl) for x in pyrs[:k + 1]], poses[:k + 1], est, cam, depth, edges)
        assert list(q.histogram) == o["histogram"] and list(q.overlaps) == o["overlaps"], (k, list(q.histogram), o)
        assert q.out_of_bounds == o["out_of_bounds"] and q.n_frames == o["n_frames"] == k + 1
        assert abs(q.overlap_measure - o["overlap_measure"]) < 1e-3 and status == o["status"]
    # a fourth frame: only the first nFramesHistogramVoting of the list vote until clearUpPastLists() trims it (tracker.cpp:143)
    trk.addOldPclAndPose(pyrs[0], poses[1], 3.0)
    trk.assessTrackingQuality(est, cur)
    assert trk.last_quality.n_frames == 3 and list(trk.last_quality.histogram) == o["histogram"]
    trk.clearUpPastLists()
    assert len(trk.mPastPcl) == 3 and trk.mPastPcl[0][2] == 1.0


def test_config3_1280x960_five_levels(ctx, orc32, orc64, engine):
    """BASELINE.json configs[2]: one 1280x960 pair, 5-level pyramid, Huber weights (always on, optimizer.h:75,156-160).
    Pyramids built by the CUDA path are bit-exact against the oracle on all five levels (level 3 and 4 never run the edge
    fill-in) and the full coarse-to-fine track agrees with the oracle like at VGA."""
    from oracle import oracle as O
    from revo_b200 import api

    p = synth_pair(3, 1280, 960)
    st = _settings(p["cam"], 5)
    gk = api.ImgPyramidRGBD(ctx, st, None, *p["key"])
    gc = api.ImgPyramidRGBD(ctx, st, None, *p["cur"])
    gk.makeKeyframe()
    cfg = O.PyrCfg(n_levels=5)
    ok = O.build_pyramid(orc64, cfg, p["cam"], *p["key"], backend="cv2")
    O.make_keyframe(orc64, ok, backend="cv2")
    oc = O.build_pyramid(orc64, cfg, p["cam"], *p["cur"], backend="cv2")
    for l in range(5):
        assert np.array_equal(gk.returnEdges(l), ok.edges[l]), f"edges L{l}"
        assert np.array_equal(gk.returnDistTransform(l), ok.dt[l]), f"dt L{l}"
        assert np.array_equal(gc.return3DEdges(l), oc.edges3d[l]), f"edges3d L{l}"
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    status, Rg, Tg, err = trk.trackFrames(np.eye(3), np.zeros(3), gk, gc)
    r32 = orc32.track_frames(ok, oc, np.eye(3), np.zeros(3), orc32.default_cfg(), 4, 0, True)
    r64 = orc64.track_frames(ok, oc, np.eye(3), np.zeros(3), orc64.default_cfg(), 4, 0, True)
    spread_r = rot_angle(r32["R"], r64["R"])
    spread_t = np.linalg.norm(r32["T"].astype(np.float64) - r64["T"])
    d_r = min(rot_angle(Rg, r32["R"]), rot_angle(Rg, r64["R"]))
    d_t = min(np.linalg.norm(Tg - r32["T"]), np.linalg.norm(Tg - r64["T"]))
    Tgt = p["T_kf_cur"]
    print(f"1280x960: gpu evals {list(trk.last_result.n_evals)[:5]} f64 {r64['evals'][:5]} d_r {d_r:.2e} d_t {d_t:.2e} "
          f"spread {spread_r:.2e}/{spread_t:.2e} gt {rot_angle(Rg, Tgt[:3, :3]):.2e}/{np.linalg.norm(Tg - Tgt[:3, 3]):.2e}")
    assert status == r64["status"]
    assert d_r <= max(2e-4, 1.5 * spread_r) and d_t <= max(2e-4, 1.5 * spread_t)
    assert rot_angle(Rg, Tgt[:3, :3]) <= rot_angle(r64["R"], Tgt[:3, :3]) + 3e-4
    assert np.linalg.norm(Tg - Tgt[:3, 3]) <= np.linalg.norm(r64["T"] - Tgt[:3, 3]) + 1e-3
    assert list(trk.last_result.n_pts)[:5] == [len(oc.edges3d[l]) for l in range(5)]


def test_revo_main_loop_on_gpu(ctx, orc32, engine):
    """revo_b200/system.py (REVO::start mirror: motion model, quality vote, keyframe switch) driving the CUDA classes over a
    short synthetic stream: follows the ground truth and agrees with the same loop over the oracle-backed stand-ins."""
    from _oracle_system import OraclePyr, OracleTracker
    from oracle import oracle as O
    from revo_b200 import api, synth
    from revo_b200.system import REVO

    w, h, n = 320, 240, 7
    s = synth.make_stream(77, n, w, h, max_trans=0.03, max_rot_deg=1.5)
    cam = synth.intrinsics(w, h)
    st = _settings(cam, 3)
    g = REVO(api.TrackerNew(ctx, api.TrackerSettings(), st))
    o = REVO(OracleTracker(orc32, 3))
    cfg = O.PyrCfg(n_levels=3)
    for i in range(n):
        bgr, depth = s["frames"][i]
        g.processFrame(api.ImgPyramidRGBD(ctx, st, None, bgr, depth, 0.033 * i))
        o.processFrame(OraclePyr(orc32, cfg, cam, bgr, depth, 0.033 * i))
    tg, to = g.trajectory(), o.trajectory()
    assert g.retracked == o.retracked and g.nKeyFrames == o.nKeyFrames
    for i in range(1, n):
        D = np.linalg.inv(to[i].astype(np.float64)) @ tg[i].astype(np.float64)
        assert rot_angle(np.eye(3), D[:3, :3]) < 1e-3 and np.linalg.norm(D[:3, 3]) < 2e-3, (i, D)
        T_gt = np.linalg.inv(s["T_w_c"][0]) @ s["T_w_c"][i]
        E = np.linalg.inv(T_gt) @ tg[i].astype(np.float64)
        assert rot_angle(np.eye(3), E[:3, :3]) < 8e-3 and np.linalg.norm(E[:3, 3]) < 2e-2
    q = g.mTracker.last_quality
    assert q is not None and q.n_frames == 3 and sum(q.histogram) > 0


def test_multi_stream_main_loop_on_gpu(ctx, engine):
    """MultiStreamREVO over the CUDA classes (one trackFramesBatch launch for all streams, a second one for the re-tracks)
    against separate single-stream REVO runs on the same device: identical trajectories and keyframe decisions."""
    from revo_b200 import api, synth
    from revo_b200.system import REVO, MultiStreamREVO, cuda_track_batch

    w, h, n, B = 320, 240, 7, 3
    cam = synth.intrinsics(w, h)
    st = _settings(cam, 3)
    streams = [synth.make_stream(300 + b, n, w, h, max_trans=0.01 + 0.04 * (b == 1), max_rot_deg=0.5 + 2.5 * (b == 1)) for b in range(B)]

    def pyr(b, i):
        return api.ImgPyramidRGBD(ctx, st, None, *streams[b]["frames"][i], 0.033 * i)

    single = [REVO(api.TrackerNew(ctx, api.TrackerSettings(), st)) for _ in range(B)]
    for b in range(B):
        for i in range(n):
            single[b].processFrame(pyr(b, i))
    trackers = [api.TrackerNew(ctx, api.TrackerSettings(), st) for _ in range(B)]
    multi = MultiStreamREVO(trackers, cuda_track_batch(trackers[0]))
    for i in range(n):
        multi.processFrames([pyr(b, i) for b in range(B)])
    for b in range(B):
        assert np.array_equal(multi.trajectories()[b], single[b].trajectory())
        assert multi.streams[b].retracked == single[b].retracked
    assert multi.batch_sizes.count(B) >= n - 1


def test_batched_vote_and_point_list_copies(ctx, engine):
    """revo_track_quality_batch == the single-vote entry stream by stream; revo_pyr_copy_points_batch copies exactly the
    level-2 lists, the copies outlive their frames, and what a points-only handle cannot do is refused (not crashed on)."""
    import ctypes as C

    from revo_b200 import api, synth

    w, h, B, n = 320, 240, 5, 5
    cam = synth.intrinsics(w, h)
    st = _settings(cam, 3)
    streams = [synth.make_stream(500 + b, n, w, h, max_trans=0.02, max_rot_deg=1.0) for b in range(B)]
    trackers = [api.TrackerNew(ctx, api.TrackerSettings(), st) for _ in range(B)]
    lists_ref = []
    for i in range(n - 1):
        pyrs = [api.ImgPyramidRGBD(ctx, st, None, *streams[b]["frames"][i], 0.033 * i) for b in range(B)]
        pls = api.copy_point_lists(ctx, pyrs, 2)
        for b in range(B):
            if i == 0:
                lists_ref.append((pls[b], pyrs[b].return3DEdgesDeviceOrder(2)))
            if b != 3 or i < 2:            # stream 3 votes with two past frames only
                trackers[b].addOldPclAndPose(pls[b], streams[b]["T_w_c"][i].astype(np.float32), 0.033 * i)
        del pyrs                           # the frames go away; the copies stay
    for pl, want in lists_ref:
        assert np.array_equal(pl.download(2), want)
        with pytest.raises(api.RevoError):
            pl.download(1)
    cur = [api.ImgPyramidRGBD(ctx, st, None, *streams[b]["frames"][n - 1], 0.0) for b in range(B)]
    est = [streams[b]["T_w_c"][n - 1].astype(np.float32) for b in range(B)]
    single = []
    for b in range(B):
        single.append((trackers[b].assessTrackingQuality(est[b], cur[b]), trackers[b].last_quality))
    status = api.assess_tracking_quality_batch(trackers, est, cur)
    for b in range(B):
        q, q1 = trackers[b].last_quality, single[b][1]
        assert status[b] == single[b][0]
        assert list(q.histogram) == list(q1.histogram) and list(q.overlaps) == list(q1.overlaps)
        assert q.out_of_bounds == q1.out_of_bounds and q.n_frames == q1.n_frames == (2 if b == 3 else 3)
        assert sum(q.histogram) > 0 and sum(q.histogram[1:]) > 0
    # a points-only handle is not a frame: no keyframe promotion, no tracking, no vote as the current frame
    pl = lists_ref[0][0]
    arr = (C.c_void_p * 1)(pl.h)
    assert ctx.lib.revo_pyr_make_keyframe_batch(ctx.h, 1, arr) == api.REVO_ERR_UNSUPPORTED
    res = api.revo_quality_result()
    e = np.eye(4, dtype=np.float32)
    assert ctx.lib.revo_track_quality(ctx.h, pl.h, 2, 0, None, None, e.ctypes.data, 3, C.byref(res)) == api.REVO_ERR_UNSUPPORTED


def test_stream_tracker_vote_policy_on_gpu(ctx, engine):
    """StreamTracker(kf_policy="vote") over the CUDA backend (batched alignment, batched vote, batched promotions, handle
    arrays) against separate REVO main loops on the same device: same keyframe decisions and world poses; every frame
    handle is released at the end."""
    from revo_b200 import api, synth
    from revo_b200.stream import CudaBackend, StreamTracker
    from revo_b200.system import REVO

    w, h, n, B = 320, 240, 9, 4
    cam = synth.intrinsics(w, h)
    st = _settings(cam, 3)
    streams = [synth.make_stream(300 + b, n, w, h, max_trans=0.01 + 0.04 * (b % 2), max_rot_deg=0.5 + 2.5 * (b % 2)) for b in range(B)]
    single = [REVO(api.TrackerNew(ctx, api.TrackerSettings(), st)) for _ in range(B)]
    for b in range(B):
        for i in range(n):
            single[b].processFrame(api.ImgPyramidRGBD(ctx, st, None, *streams[b]["frames"][i], 0.033 * i))
    be = CudaBackend(ctx, st)
    trk = StreamTracker(be, B, kf_policy="vote")
    trk.keep_history = True
    frames = lambda i: (np.stack([streams[b]["frames"][i][0] for b in range(B)]), np.stack([streams[b]["frames"][i][1] for b in range(B)]))
    trk.start(*frames(0))
    flags = []
    for i in range(1, n):
        trk.step(*frames(i))
        flags.append(trk.just_added.copy())
    assert trk.n_retracks == sum(len(s.retracked) for s in single) >= 1
    for b in range(B):
        assert [i + 1 for i, f in enumerate(flags) if f[b]] == single[b].retracked
        traj = single[b].trajectory()
        for i in range(1, n):
            assert np.array_equal(trk.history[i - 1][0][b], traj[i]), (b, i)
    trk.close()
