"""The oracle against the REFERENCE ITSELF: ``oracle/_ref/librevo_ref.so`` is built from the reference's own sources
(``datastructures/imgpyramidrgbd.cpp``, ``system/optimizer.cpp``, ``system/tracker.cpp``, ``utils/LGSX.h``, ... compiled verbatim
where they lie under /root/reference against the API shims of ``oracle/shim/``; recipe: ``make -C oracle ref``).  Every result
of the restatement in ``oracle/revo_oracle.c`` / ``oracle/oracle.py`` that the GPU parity tests rely on must equal what the
reference's code computes on the same input -- bit for bit for the pyramid arrays AND for the float32 optimizer / tracker
(same evaluation count, same LM trace, same pose).  Where /root/reference is absent (the GPU box) the prebuilt library is used
if it travelled along, else these tests skip and ``tests/test_oracle.py`` checks the oracle against the committed golden
vectors generated from the same library (``tests/golden/make_ref_golden.py``)."""
import numpy as np
import pytest

from conftest import synth_pair
from oracle import oracle as O
from oracle import ref as RF

pytestmark = pytest.mark.skipif(not RF.build(), reason="oracle/_ref is not built and /root/reference is absent")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def _pyramids(orc32, seed, w, h, n_percentage=0.3, depth_holes=True):
    p = synth_pair(seed, w, h)
    (kb, kd), (cb, cd) = p["key"], p["cur"]
    if depth_holes:
        kd, cd = kd.copy(), cd.copy()
        kd[5:15, 20:50] = np.nan
        cd[h // 2, :] = 0.0
    cfg = O.PyrCfg(n_levels=3)
    cfg.n_percentage = n_percentage
    ok = O.build_pyramid(orc32, cfg, p["cam"], kb, kd)
    O.make_keyframe(orc32, ok)
    oc = O.build_pyramid(orc32, cfg, p["cam"], cb, cd)
    rk = RF.RefPyramid(p["cam"], 3, kb, kd, n_percentage=n_percentage)
    rk.make_keyframe()
    rc = RF.RefPyramid(p["cam"], 3, cb, cd, n_percentage=n_percentage)
    return p, ok, oc, rk, rc


@pytest.mark.parametrize("seed,w,h,n_percentage", [(1, 320, 240, 0.3), (4, 160, 120, 0.3), (7, 320, 240, 2.0), (2, 640, 480, 0.3)])
def test_pyramid_of_the_reference_equals_oracle(orc32, seed, w, h, n_percentage):
    """ImgPyramidRGBD(...) + makeKeyframe() (imgpyramidrgbd.cpp:43-276): every array of every level, bit for bit
    (n_percentage 2.0 forces the edge fill-in of levels 1 and 2).  Sizes are multiples of the patch size on every level, like
    all dataset configurations: elsewhere generateDistHistogram indexes its u8 matrix out of bounds (imgpyramidrgbd.cpp:153:
    column w/P wraps into the next row, row h/P lands behind the buffer), which this pin made visible on a 192x144 image."""
    p, ok, oc, rk, rc = _pyramids(orc32, seed, w, h, n_percentage)
    for l in range(3):
        for what, arr in (("gray", ok.gray[l]), ("depth", ok.depth[l]), ("edges", ok.edges[l]), ("edges_orig", ok.edges_orig[l]),
                          ("edges3d", ok.edges3d[l]), ("dt", ok.dt[l]), ("opt", ok.opt[l])):
            g = rk.get(what, l)
            assert g.shape == np.asarray(arr).shape and np.array_equal(_bits(g), _bits(np.asarray(arr, g.dtype))), (what, l)
        assert np.array_equal(_bits(rc.get("edges3d", l)), _bits(np.asarray(oc.edges3d[l], np.float32))), l
        hist = rk.get("hist", l)
        assert np.array_equal(hist, np.asarray(ok.hist[l], np.uint8).reshape(-1)), l
    if n_percentage > 1:
        assert any(not np.array_equal(ok.edges[l], ok.edges_orig[l]) for l in (1, 2)), "fill-in did not fire"


@pytest.mark.parametrize("seed", [1, 3, 22])
def test_optimizer_of_the_reference_equals_oracle(orc32, seed):
    """Optimizer::trackFrames (optimizer.cpp:235-311) level by level, coarse to fine: same number of evaluations, same LM
    trace (error of every try), same pose and residual info -- bit for bit; and one PASS A + PASS B evaluation
    (calcErrorAndBuffers + calculateWarpUpdate + LGS6) against the record of the oracle."""
    p, ok, oc, rk, rc = _pyramids(orc32, seed, 320, 240)
    ocfg = orc32.default_cfg()
    R, T = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    for lvl in (2, 1, 0):
        r = RF.opt_track_level(rk, rc, ocfg, lvl, R, T)
        q = orc32.track_level(oc.edges3d[lvl], ok.opt[lvl], oc.cams[lvl], R, T, ocfg, lvl)
        assert r["n_evals"] == q["n_evals"], (lvl, r["n_evals"], q["n_evals"])
        assert [e for _, _, e in r["trace"]] == pytest.approx([t[0] for t in q["trace"]], abs=6e-7)      # the log prints 6 decimals
        assert [(g, b) for g, b, _ in r["trace"]] == [(t[3], t[4]) for t in q["trace"]]
        assert np.array_equal(r["R"], q["R"].astype(np.float32)) and np.array_equal(r["T"], q["T"].astype(np.float32)), lvl
        assert np.float32(r["error"]) == np.float32(q["error"]) and (r["good"], r["bad"]) == (q["good"], q["bad"])
        assert np.float32(r["sum_w"]) == np.float32(q["sum_w"]) and np.float32(r["sum_unw"]) == np.float32(q["sum_unw"])
        R, T = r["R"], r["T"]
        # one evaluation at the level's result pose
        e = RF.opt_eval(rk, rc, ocfg, lvl, R, T)
        rec = orc32.eval_record(oc.edges3d[lvl], ok.opt[lvl], oc.cams[lvl], R, T, ocfg, lvl)
        n = rec[29]
        assert (e["good"], e["bad"]) == (int(rec[29]), int(rec[30]))
        A = np.zeros((6, 6), np.float32)
        k = 0
        for i in range(6):
            for j in range(i, 6):
                A[i, j] = A[j, i] = np.float32(rec[k] / n)
                k += 1
        assert np.array_equal(A, e["A"]), lvl                                           # LGS6::finish: A / n
        assert np.array_equal((-rec[21:27] / n).astype(np.float32), e["b"]), lvl        # ls.b is the NEGATIVE sum / n
        assert np.float32(rec[27]) == np.float32(e["sum_w"]) and np.float32(rec[28]) == np.float32(e["sum_unw"])
    # edge filter off (the OptimizerSettings default)
    ocfg.use_edge_filter = 0
    r = RF.opt_track_level(rk, rc, ocfg, 1, np.eye(3), np.zeros(3))
    q = orc32.track_level(oc.edges3d[1], ok.opt[1], oc.cams[1], np.eye(3), np.zeros(3), ocfg, 1)
    assert r["n_evals"] == q["n_evals"] and np.array_equal(r["R"], q["R"].astype(np.float32)) and np.array_equal(r["T"], q["T"].astype(np.float32))


@pytest.mark.parametrize("seed,w,h", [(1, 320, 240), (23, 320, 240), (2, 640, 480)])
def test_tracker_of_the_reference_equals_oracle(orc32, seed, w, h):
    """TrackerNew::trackFrames with checkInitializationValues / evalCostFunction (tracker.cpp:265-393); the last case is the
    benchmark's frame size (3 levels: the compiled reference divides by zero at a fourth)."""
    from revo_b200 import synth

    p, ok, oc, rk, rc = _pyramids(orc32, seed, w, h)
    ocfg = orc32.default_cfg()
    trk = RF.RefTracker(rk, ocfg)
    bad = synth.se3_exp([0.4, 0.3, -0.2, 0.2, -0.15, 0.1])
    near = synth.se3_exp([0.004, -0.002, 0.003, 0.002, -0.001, 0.0015])
    for M in (np.eye(4), near, bad):
        R0, T0 = np.asarray(M[:3, :3], np.float32), np.asarray(M[:3, 3], np.float32)
        r = trk.track_frames(rk, rc, R0, T0)
        q = orc32.track_frames(ok, oc, R0, T0, ocfg, 2, 0, True)
        assert r["status"] == q["status"]
        assert np.array_equal(r["R"], q["R"].astype(np.float32)) and np.array_equal(r["T"], q["T"].astype(np.float32))
        assert np.float32(r["error"]) == np.float32(q["error"])
        assert len(r["trace"]) == sum(q["evals"][:3]) - 3
        c_ref = trk.eval_cost(rk, rc, R0, T0, 2)
        c_orc = orc32.eval_cost_function(oc.edges3d[2], ok.dt[2], oc.cams[2], R0, T0, ocfg, 2)
        assert np.float32(c_ref) == np.float32(c_orc)


def test_quality_vote_of_the_reference_equals_oracle(orc32):
    """TrackerNew::assessTrackingQuality / addOldPclAndPose / clearUpPastLists (tracker.cpp:118-257) against the numpy
    restatement the GPU vote kernel is tested with: status and the histogram / overlap counts the reference logs."""
    from revo_b200 import synth

    seeds = (1, 23, 24, 25)
    built = [_pyramids(orc32, s, 320, 240, depth_holes=False) for s in seeds]
    p0, ok0, oc0, rk0, rc0 = built[0]
    trk = RF.RefTracker(rk0, orc32.default_cfg())
    rng = np.random.default_rng(5)
    poses = [synth.se3_exp(rng.normal(0, 0.01, 6)) for _ in seeds]
    est = synth.se3_exp(rng.normal(0, 0.01, 6))
    lvl = 2
    cam = (oc0.cams[lvl].fx, oc0.cams[lvl].fy, oc0.cams[lvl].cx, oc0.cams[lvl].cy, oc0.cams[lvl].w, oc0.cams[lvl].h)
    assert trk.assess(rc0, est)["status"] == 0          # nothing to vote with
    for k in range(len(seeds)):
        trk.add_old(built[k][3], poses[k], float(k))
        r = trk.assess(rc0, est)
        o = O.assess_tracking_quality([b[1].edges3d[lvl] for b in built[:k + 1]], poses[:k + 1], est, cam, oc0.depth[lvl],
                                      oc0.edges_orig[lvl])
        nf = o["n_frames"]
        assert r["status"] == o["status"], (k, r, o)
        assert r["histogram"] == o["histogram"][:nf + 1] and r["overlaps"] == o["overlaps"][:nf + 1], (k, r, o)
        assert r["out_of_bounds"] == o["out_of_bounds"]
    assert trk.num_past() == 4
    trk.clear_past()
    assert trk.num_past() == 3


def test_lgs6_and_interpolation_of_the_reference(orc32):
    """LGS6::initialize / update / finish (LGSX.h:196-204,392-398,320-326) and getInterpolatedElement43 (optimizer.h:173-185)
    against float32 numpy in the reference's operation order."""
    rng = np.random.default_rng(3)
    n = 257
    J = rng.normal(0, 1, (n, 6)).astype(np.float32)
    res = rng.uniform(0, 3, n).astype(np.float32)
    w = rng.uniform(0.1, 1, n).astype(np.float32)
    A, b, err = RF.lgs6(J, res, w, finish=True)
    A0, b0, e0 = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), np.float32(0)
    for i in range(n):
        A0 += (np.outer(J[i], J[i]).astype(np.float32) * w[i]).astype(np.float32)      # (J J^T) * w
        b0 -= (J[i] * np.float32(res[i] * w[i])).astype(np.float32)                     # J * (res * w)
        e0 += np.float32(np.float32(res[i] * res[i]) * w[i])
    assert np.array_equal(A, A0 / np.float32(n)) and np.array_equal(b, b0 / np.float32(n)) and np.float32(err) == e0 / np.float32(n)
    opt = rng.normal(0, 1, (12, 16, 4)).astype(np.float32)
    for x, y in ((3.25, 4.75), (7.0, 2.5), (10.999, 9.001)):
        ix, iy = int(x), int(y)
        dx, dy = np.float32(x) - np.float32(ix), np.float32(y) - np.float32(iy)
        dxdy = dx * dy
        want = (dxdy * opt[iy + 1, ix + 1, :3] + (dy - dxdy) * opt[iy + 1, ix, :3] + (dx - dxdy) * opt[iy, ix + 1, :3]
                + (np.float32(1) - dx - dy + dxdy) * opt[iy, ix, :3])
        assert np.array_equal(RF.interp43(opt, x, y), want.astype(np.float32))


def test_colored_point_cloud_of_the_reference(orc32):
    """generateColoredPcl (imgpyramidrgbd.cpp:279-327), edge cloud and dense cloud, against the restatement the device kernel is
    tested with (revo_b200.api.colored_pcl_from_arrays)."""
    import cv2

    from revo_b200.api import colored_pcl_from_arrays

    p, ok, oc, rk, rc = _pyramids(orc32, 1, 320, 240)
    bgr = p["key"][0]
    for lvl in range(3):
        rgb = bgr
        for _ in range(lvl):
            rgb = cv2.pyrDown(rgb)
        c = ok.cams[lvl]
        for dense in (False, True):
            want = colored_pcl_from_arrays(rgb, ok.depth[lvl], ok.edges[lvl], c.fx, c.fy, c.cx, c.cy, 0.1, 5.2, dense)
            got = rk.colored_pcl(lvl, dense)
            assert got.shape == want.shape, (lvl, dense)
            # the reference sizes the matrix cam.area / 5 for the edge cloud (imgpyramidrgbd.cpp:283) and writes past its end when
            # more than a fifth of the pixels are edge points (the coarse levels): only the columns that fit are defined
            fit = want.shape[1] if dense else min(want.shape[1], int(c.w * c.h / 5.0))
            assert np.array_equal(_bits(got[:, :fit]), _bits(want[:, :fit])), (lvl, dense)
