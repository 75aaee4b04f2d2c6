"""The tracking KERNEL itself on the CPU: ``k_track`` (track.cu) -- point cache, software-pipelined gather loop, transposing
warp-shuffle reduction, per-level LM loop, result record -- compiled from its CUDA source text against the emulation layer of
``tests/_cuda_emu.py`` (one OS thread per CUDA thread, clusters of one CTA) and compared with the float64 oracle after the same
number of LM tries: the `-m gpu` test ``test_track_level_fixed_iterations`` without a GPU.  The multi-CTA exchange and the
multi-GPU mailboxes are not emulated (hardware paths)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rot_angle, synth_pair

# an emulation deadlock must not hang the suite (the C call cannot be interrupted by a signal: kill the run instead)
pytestmark = pytest.mark.timeout(900, method="thread")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    import _cuda_emu

    return _cuda_emu.build(str(tmp_path_factory.mktemp("cuda_emu")))


def run_kernel(lib, variant, pairs, cfg, mode=0, level=0, n_ctas=1, pcap=18, ctas_per_pair=1):
    """pairs: list of (kf oracle pyramid, cur oracle pyramid, R (3,3), T (3,)).  Returns the revo_track_result records."""
    from revo_b200 import api

    n, NL = len(pairs), pairs[0][0].n_levels
    keep, pts, npts, dts, ws, hs, cams = [], [], [], [], [], [], []
    for kf, cur, _, _ in pairs:
        for l in range(NL):
            p = np.ascontiguousarray(cur.edges3d[l], np.float32)
            d = np.ascontiguousarray(kf.dt[l], np.float32)
            keep += [p, d]
            pts.append(p.ctypes.data); npts.append(len(p)); dts.append(d.ctypes.data)
            c = cur.cams[l]
            ws.append(c.w); hs.append(c.h); cams += [c.fx, c.fy, c.cx, c.cy]
    arr = lambda v, t: np.ascontiguousarray(v, t)      # noqa: E731
    pts_a, dts_a = (C.c_void_p * len(pts))(*pts), (C.c_void_p * len(dts))(*dts)
    npts_a, ws_a, hs_a, cams_a = arr(npts, np.int32), arr(ws, np.int32), arr(hs, np.int32), arr(cams, np.float32)
    R9 = arr(np.stack([np.asarray(R, np.float32).T.reshape(-1) for _, _, R, _ in pairs]), np.float32)
    t3 = arr(np.stack([np.asarray(T, np.float32) for _, _, _, T in pairs]), np.float32)
    out = np.zeros(n, api.TRACK_RESULT_DTYPE)
    rec = np.zeros((n, 32), np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    rc = lib.emu_track_pairs(C.c_int(variant), C.c_int(n), C.c_int(n_ctas), C.c_int(ctas_per_pair), C.c_int(NL), pts_a, vp(npts_a), dts_a, vp(ws_a), vp(hs_a),
                             vp(cams_a), vp(R9), vp(t3), C.byref(cfg), C.c_int(mode), C.c_int(level), C.c_int(pcap), vp(out), vp(rec))
    assert rc == 0
    return out, rec


def oracle_chain(orc, kf, cur, R, T, n_tries):
    """Optimizer::trackFrames level by level (coarse to fine) with a fixed number of LM tries."""
    evals = []
    for lvl in range(kf.n_levels - 1, -1, -1):
        ocfg = orc.default_cfg()
        for l in range(6):
            ocfg.convergence_eps[l] = 2.0
        r = orc.track_level(cur.edges3d[lvl], kf.opt[lvl], cur.cams[lvl], R, T, ocfg, lvl, max_tries=n_tries)
        R, T = r["R"].astype(np.float32), r["T"].astype(np.float32)
        evals.append(r["n_evals"])
    return R, T, evals[::-1], r


def build_pair(orc, seed, w=320, h=240, n_levels=3):
    from oracle import oracle as O

    p = synth_pair(seed, w, h)
    cfg = O.PyrCfg(n_levels=n_levels)
    kf = O.build_pyramid(orc, cfg, p["cam"], *p["key"])
    O.make_keyframe(orc, kf)
    cur = O.build_pyramid(orc, cfg, p["cam"], *p["cur"])
    return kf, cur


def tracker_cfg(n_tries, n_levels=3, check_init=0):
    from revo_b200 import api

    cfg = api.revo_tracker_config()
    cfg.check_init_values = check_init
    cfg.pyr_min_lvl, cfg.pyr_max_lvl = n_levels - 1, 0
    cfg.opt = api.OptimizerSettings(USE_EDGE_FILTER=True, max_lm_tries=n_tries, convergenceEps=[2.0] * 6)._c()
    return cfg


def test_k_track_on_host_matches_oracle_after_same_iterations(emu, orc64):
    from revo_b200 import synth

    T0 = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
    R0, t0 = np.asarray(T0[:3, :3], np.float32), np.asarray(T0[:3, 3], np.float32)
    pairs = [build_pair(orc64, seed) + (R0, t0) for seed in (1, 22)]
    n_tries = 5
    # two pairs, one CTA: the persistent loop fetches the second pair from the work counter; pcap 2 forces the uncached tail
    for n_ctas, pcap in ((1, 18), (2, 2)):
        out, _ = run_kernel(emu, 0, pairs, tracker_cfg(n_tries), n_ctas=n_ctas, pcap=pcap)
        for i, (kf, cur, _, _) in enumerate(pairs):
            Ro, To, evals, last = oracle_chain(orc64, kf, cur, R0, t0, n_tries)
            R = out["R"][i].reshape(3, 3).T
            assert out["rc"][i] == 0 and list(out["n_evals"][i][:3]) == evals, (out["n_evals"][i], evals)
            assert rot_angle(R, Ro) <= 1e-4 and np.linalg.norm(out["t"][i] - To) <= 1e-4
            assert abs(out["error"][i] - last["error"]) <= 1e-4 * max(1.0, abs(last["error"]))
            assert out["good"][i] == last["good"] and out["bad"][i] == last["bad"]
            assert list(out["n_pts"][i][:3]) == [len(cur.edges3d[l]) for l in range(3)]


def test_k_track_on_host_single_evaluation_and_init_check(emu, orc64):
    """mode 2 (one fused evaluation, the record export of revo_eval) and the identity-vs-initial-pose check of mode 0."""
    from revo_b200 import synth

    kf, cur = build_pair(orc64, 3)
    near = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
    R, T = np.asarray(near[:3, :3], np.float32), np.asarray(near[:3, 3], np.float32)
    ocfg = orc64.default_cfg()
    for lvl in range(3):
        _, rec = run_kernel(emu, 0, [(kf, cur, R, T)], tracker_cfg(0), mode=2, level=lvl)
        o = orc64.eval_record(cur.edges3d[lvl], kf.opt[lvl], cur.cams[lvl], R, T, ocfg, lvl)
        assert abs(rec[0][29] - o[29]) <= 1 and rec[0][29] + rec[0][30] == o[29] + o[30]
        sA = np.abs(o[:21]).max()
        assert np.abs(rec[0][:21] - o[:21]).max() <= 3e-5 * sA and abs(rec[0][27] - o[27]) <= 3e-5 * abs(o[27])
    # checkInitializationValues (tracker.cpp:265-283): the pose is reset to the identity when the identity costs less.  The
    # decision is the FLOAT32 reference's: at the identity every point projects onto a pixel corner, where float32 and
    # float64 floor differently (for this pair the float64 oracle would keep the true pose, the float one resets it).
    from oracle import oracle as O

    orc32 = O.Oracle("f32")
    o32 = orc32.default_cfg()
    pts, dtm, cam = cur.edges3d[2], kf.dt[2], cur.cams[2]
    cost_eye = orc32.eval_cost_function(pts, dtm, cam, np.eye(3), np.zeros(3), o32, 2)
    M = synth_pair(3, 320, 240)["T_kf_cur"]
    decisions = []
    for P in (synth.se3_exp([0.3, -0.2, 0.1, 0.1, 0.2, -0.1]), M, near):
        Rp, Tp = np.asarray(P[:3, :3], np.float32), np.asarray(P[:3, 3], np.float32)
        want = cost_eye < orc32.eval_cost_function(pts, dtm, cam, Rp, Tp, o32, 2)
        out, _ = run_kernel(emu, 0, [(kf, cur, Rp, Tp)], tracker_cfg(3, check_init=1))
        assert out["rc"][0] == 0 and bool(out["used_identity_init"][0]) == want
        decisions.append(want)
    assert decisions[0]                                          # a wildly wrong pose always loses against the identity
    out, _ = run_kernel(emu, 0, [(kf, cur, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))], tracker_cfg(3, check_init=1))
    assert out["used_identity_init"][0] == 0                     # equal costs: the initial pose stays (strict <, :277)
    # not a rotation: the pair is refused with the error code that replaces Sophus' abort()
    out, _ = run_kernel(emu, 0, [(kf, cur, 1.01 * np.eye(3, dtype=np.float32), T)], tracker_cfg(3))
    assert out["rc"][0] == 5


@pytest.mark.parametrize("ctas_per_pair", [2, 4])
def test_k_track_on_host_multi_cta_cluster_exchange(emu, orc64, ctas_per_pair):
    """Clusters of several CTAs: the block-cyclic split of the point list over the CTAs and the one-sided exchange of the
    per-CTA partials (st.async into every peer's shared memory + transaction barrier, emulated) must give the oracle's result;
    two pairs on one cluster exercise the hand-over of the next pair through distributed shared memory."""
    from revo_b200 import synth

    T0 = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
    R0, t0 = np.asarray(T0[:3, :3], np.float32), np.asarray(T0[:3, 3], np.float32)
    pairs = [build_pair(orc64, seed) + (R0, t0) for seed in (1, 22)]
    n_tries = 4
    out, _ = run_kernel(emu, 0, pairs, tracker_cfg(n_tries), n_ctas=1, pcap=4, ctas_per_pair=ctas_per_pair)
    for i, (kf, cur, _, _) in enumerate(pairs):
        Ro, To, evals, last = oracle_chain(orc64, kf, cur, R0, t0, n_tries)
        R = out["R"][i].reshape(3, 3).T
        assert out["rc"][i] == 0 and list(out["n_evals"][i][:3]) == evals
        assert rot_angle(R, Ro) <= 1e-4 and np.linalg.norm(out["t"][i] - To) <= 1e-4
        assert out["good"][i] == last["good"] and out["bad"][i] == last["bad"]
