"""GPU parity on a POPULATION of frame pairs ("what's weak" 1-2): the whole CUDA path -- pyramids built
on the device, keyframe promotion, TrackerNew::trackFrames with the reference's own termination rules -- against the
float32 "reference-as-is" oracle and the float64 oracle on >= 32 VGA pairs with 4 levels.

What is asserted (tolerance of the path: 1e-4 rad / 1e-4 m after the same number of iterations, BASELINE.json north_star):
 * every pair whose LM trace (accept / reject sequence of every level) equals the float64 oracle's: pose within 1e-4 rad /
   1e-4 m of it; same trace as the float32 oracle: within 1e-4 or that oracle's own distance from the float64 one;
 * fixed-iteration mode (same number of LM tries on both sides by construction, all levels chained): EVERY pair within
   1e-4 rad / 1e-4 m of the float64 oracle, same evaluation counts, same good/bad counts.
What is reported, not hidden: the pairs whose default-rule traces differ (the accept / convergence tests of
optimizer.cpp:273-279 sit on float-rounding knife edges; the two oracle precisions disagree with each other the same way),
with their pose distance.  The summary is written to gpurun_out/parity_population.json (copied under profiles/ by hand).
"""
import json
import os

import numpy as np
import pytest

from conftest import rot_angle

pytestmark = pytest.mark.gpu

N_PAIRS = 32
LEVELS = 4


@pytest.fixture(scope="module")
def population(ctx, orc32, orc64):
    """32 VGA pairs (frame 0 = keyframe, frame 2 = tracked frame of 32 distinct synthetic streams), pyramids on both sides."""
    import torch

    from oracle import oracle as O
    from revo_b200 import api, synth_torch

    w, h, gap = 640, 480, 2
    dev = torch.device("cuda", 0)
    bgr = torch.empty((gap + 1, N_PAIRS, h, w, 3), dtype=torch.uint8)
    depth = torch.empty((gap + 1, N_PAIRS, h, w), dtype=torch.float32)
    cam, poses = synth_torch.render_streams([7000 + s for s in range(N_PAIRS)], gap + 1, w, h, dev, bgr, depth)
    bgr, depth = bgr.numpy(), depth.numpy()
    fx, fy, cx, cy, _, _ = cam
    st = api.ImgPyramidSettings(PYR_MIN_LVL=LEVELS - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    kf = api.PyramidBatch(ctx, st, np.ascontiguousarray(bgr[0]), np.ascontiguousarray(depth[0]), N_PAIRS)
    cur = api.PyramidBatch(ctx, st, np.ascontiguousarray(bgr[gap]), np.ascontiguousarray(depth[gap]), N_PAIRS)
    kf.makeKeyframes()
    ctx.synchronize()
    cfg = O.PyrCfg(n_levels=LEVELS)
    oks, ocs = [], []
    for i in range(N_PAIRS):
        ok = O.build_pyramid(orc32, cfg, cam, bgr[0, i], depth[0, i])
        O.make_keyframe(orc32, ok)
        oks.append(ok)
        ocs.append(O.build_pyramid(orc32, cfg, cam, bgr[gap, i], depth[gap, i]))
    T_gt = [np.linalg.inv(poses[i][0]) @ poses[i][gap] for i in range(N_PAIRS)]
    yield dict(st=st, kf=kf, cur=cur, oks=oks, ocs=ocs, T_gt=T_gt, cam=cam)
    kf.destroy()
    cur.destroy()


def test_population_pyramids_bit_exact(ctx, population):
    """The inputs of the tracker are the same on both sides: 3-D lists and lookup structures of all 32 pairs."""
    P = population
    for i in range(0, N_PAIRS, 5):
        for l in range(LEVELS):
            assert np.array_equal(P["cur"][i].return3DEdges(l), P["ocs"][i].edges3d[l]), (i, l)
            assert np.array_equal(P["kf"][i].returnOptimizationStructure(l), P["oks"][i].opt[l]), (i, l)


def test_population_default_rules(ctx, orc32, orc64, population):
    from revo_b200 import api

    P = population
    trk = api.TrackerNew(ctx, api.TrackerSettings(), P["st"])
    I = np.tile(np.eye(3, dtype=np.float32), (N_PAIRS, 1, 1))
    Z = np.zeros((N_PAIRS, 3), np.float32)
    out, traces = trk.trackFramesBatch(I, Z, P["kf"], P["cur"], trace_cap=512)
    rows = []
    for i in range(N_PAIRS):
        r32 = orc32.track_frames_traced(P["oks"][i], P["ocs"][i], np.eye(3), np.zeros(3), orc32.default_cfg(), LEVELS - 1, 0, True)
        r64 = orc64.track_frames_traced(P["oks"][i], P["ocs"][i], np.eye(3), np.zeros(3), orc64.default_cfg(), LEVELS - 1, 0, True)
        Rg, Tg = api.result_R(out[i]), out[i]["t"].astype(np.float64)
        ev = [int(x) for x in out[i]["n_evals"][:LEVELS]]
        acc_gpu = {l: "".join("A" if t[2] else "r" for t in traces[i] if t[5] == l) for l in range(LEVELS)}
        Tgt = P["T_gt"][i]
        rows.append(dict(
            pair=i, evals_gpu=ev, evals_f32=list(r32["evals"][:LEVELS]), evals_f64=list(r64["evals"][:LEVELS]),
            same_trace_f32=all(acc_gpu[l] == r32["accepts"][l] for l in range(LEVELS)),
            same_trace_f64=all(acc_gpu[l] == r64["accepts"][l] for l in range(LEVELS)),
            oracles_same_trace=all(r32["accepts"][l] == r64["accepts"][l] for l in range(LEVELS)),
            rot_vs_f32=rot_angle(Rg, r32["R"]), trans_vs_f32=float(np.linalg.norm(Tg - r32["T"])),
            rot_vs_f64=rot_angle(Rg, r64["R"]), trans_vs_f64=float(np.linalg.norm(Tg - r64["T"])),
            rot_f32_vs_f64=rot_angle(r32["R"], r64["R"]), trans_f32_vs_f64=float(np.linalg.norm(r32["T"].astype(np.float64) - r64["T"])),
            rot_vs_gt=rot_angle(Rg, Tgt[:3, :3]), trans_vs_gt=float(np.linalg.norm(Tg - Tgt[:3, 3])),
            rot_f64_vs_gt=rot_angle(r64["R"], Tgt[:3, :3]), trans_f64_vs_gt=float(np.linalg.norm(r64["T"] - Tgt[:3, 3])),
            status_gpu=int(out[i]["status"]), status_f32=int(r32["status"]), status_f64=int(r64["status"]), rc=int(out[i]["rc"])))
    # "same trace": the same accept / reject sequence on every level (hence the same number of evaluations)
    same32 = [r for r in rows if r["same_trace_f32"]]
    same64 = [r for r in rows if r["same_trace_f64"]]
    same_oracles = [r for r in rows if r["oracles_same_trace"]]
    summary = dict(
        n_pairs=N_PAIRS, levels=LEVELS, tolerance=dict(rot_rad=1e-4, trans_m=1e-4),
        same_trace_as_f32=len(same32), same_trace_as_f64=len(same64), oracles_agree_with_each_other=len(same_oracles),
        max_rot_same_trace_f32=max([r["rot_vs_f32"] for r in same32], default=None),
        max_trans_same_trace_f32=max([r["trans_vs_f32"] for r in same32], default=None),
        max_rot_same_trace_f64=max([r["rot_vs_f64"] for r in same64], default=None),
        max_trans_same_trace_f64=max([r["trans_vs_f64"] for r in same64], default=None),
        all_pairs=dict(
            median_rot_vs_f32=float(np.median([r["rot_vs_f32"] for r in rows])), max_rot_vs_f32=max(r["rot_vs_f32"] for r in rows),
            median_trans_vs_f32=float(np.median([r["trans_vs_f32"] for r in rows])), max_trans_vs_f32=max(r["trans_vs_f32"] for r in rows),
            median_rot_vs_f64=float(np.median([r["rot_vs_f64"] for r in rows])), max_rot_vs_f64=max(r["rot_vs_f64"] for r in rows),
            median_trans_vs_f64=float(np.median([r["trans_vs_f64"] for r in rows])), max_trans_vs_f64=max(r["trans_vs_f64"] for r in rows),
            median_rot_f32_vs_f64=float(np.median([r["rot_f32_vs_f64"] for r in rows])), max_rot_f32_vs_f64=max(r["rot_f32_vs_f64"] for r in rows),
            median_trans_f32_vs_f64=float(np.median([r["trans_f32_vs_f64"] for r in rows])),
            max_trans_f32_vs_f64=max(r["trans_f32_vs_f64"] for r in rows)),
        pairs=rows)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "parity_population.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "pairs"}))
    for r in rows:
        assert r["rc"] == 0
    # the bar, wherever "the same iteration count" holds
    for r in same64:
        assert r["rot_vs_f64"] <= 1e-4 and r["trans_vs_f64"] <= 1e-4, r
    # The float32 oracle sums ~25 000 terms per level sequentially in float32 (as the reference does); the CUDA path keeps
    # float32 only inside a thread (<= ~40 terms) and sums in double above.  With the same accept / reject sequence the two
    # may therefore still differ by the float32 oracle's own accumulation error, whose size is its distance from the float64
    # oracle on that pair: the bar is 1e-4 or that distance, whichever is larger (reported per pair in the JSON).
    for r in same32:
        assert r["rot_vs_f32"] <= max(1e-4, r["rot_f32_vs_f64"]) and r["trans_vs_f32"] <= max(1e-4, r["trans_f32_vs_f64"]), r
    # Where the traces differ the runs stop at different points of the same flat minimum.  The yardstick for that is the
    # reference's own sensitivity to arithmetic precision (float32 vs float64 oracle, same algorithm, same inputs): over the
    # population the CUDA path must not be further from the float32 oracle than the float64 oracle is.
    A = summary["all_pairs"]
    assert A["median_rot_vs_f32"] <= 1.25 * A["median_rot_f32_vs_f64"] + 1e-6, A
    assert A["median_trans_vs_f32"] <= 1.25 * A["median_trans_f32_vs_f64"] + 1e-6, A
    assert A["max_rot_vs_f32"] <= 1.25 * A["max_rot_f32_vs_f64"] and A["max_trans_vs_f32"] <= 1.25 * A["max_trans_f32_vs_f64"], A
    # ... and converges to the ground truth like the oracle: mean error within 10 % of the float64 oracle's
    m_r, m_t = np.mean([r["rot_vs_gt"] for r in rows]), np.mean([r["trans_vs_gt"] for r in rows])
    o_r, o_t = np.mean([r["rot_f64_vs_gt"] for r in rows]), np.mean([r["trans_f64_vs_gt"] for r in rows])
    print(f"mean error vs ground truth: gpu {m_r:.2e} rad {m_t:.2e} m, f64 oracle {o_r:.2e} rad {o_t:.2e} m")
    assert m_r <= 1.1 * o_r + 2e-5 and m_t <= 1.1 * o_t + 5e-5
    for r in rows:
        assert r["status_gpu"] in (r["status_f32"], r["status_f64"]), r
    # the traces must coincide on part of the population, else "same trace" would be an empty promise
    assert len(same64) + len(same32) >= 3, (len(same32), len(same64))


@pytest.mark.parametrize("n_tries", [6])
def test_population_fixed_iterations(ctx, orc64, population, n_tries):
    """Same number of LM tries per level on both sides, levels chained coarse to fine: every one of the 32 pairs within
    1e-4 rad / 1e-4 m of the float64 oracle (the test that does not depend on rounding knife edges)."""
    from revo_b200 import api, synth

    P = population
    T0 = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])      # off the integer pixel grid
    worst_r = worst_t = 0.0
    osett = api.OptimizerSettings(USE_EDGE_FILTER=True, max_lm_tries=n_tries, convergenceEps=[2.0] * 6)
    tsett = api.TrackerSettings(CHECK_INIT_VALUES=False, optimizerSettings=osett)
    trk = api.TrackerNew(ctx, tsett, P["st"])
    R0 = np.tile(np.asarray(T0[:3, :3], np.float32), (N_PAIRS, 1, 1))
    t0 = np.tile(np.asarray(T0[:3, 3], np.float32), (N_PAIRS, 1))
    out = trk.trackFramesBatch(R0, t0, P["kf"], P["cur"])
    for i in range(N_PAIRS):
        Ro, To = R0[i].copy(), t0[i].copy()
        ev = []
        for lvl in range(LEVELS - 1, -1, -1):
            ocfg = orc64.default_cfg()
            for l in range(6):
                ocfg.convergence_eps[l] = 2.0
            r = orc64.track_level(P["ocs"][i].edges3d[lvl], P["oks"][i].opt[lvl], P["ocs"][i].cams[lvl], Ro, To, ocfg, lvl, max_tries=n_tries)
            Ro, To = r["R"].astype(np.float32), r["T"].astype(np.float32)
            ev.append(r["n_evals"])
        ev = ev[::-1]
        assert [int(x) for x in out[i]["n_evals"][:LEVELS]] == ev, (i, out[i]["n_evals"], ev)
        d_r, d_t = rot_angle(api.result_R(out[i]), Ro), float(np.linalg.norm(out[i]["t"] - To))
        worst_r, worst_t = max(worst_r, d_r), max(worst_t, d_t)
        assert d_r <= 1e-4 and d_t <= 1e-4, (i, d_r, d_t)
        # counts of the last evaluation: a point within rounding of the u > 1 / edge-distance thresholds may classify differently
        assert int(out[i]["good"]) + int(out[i]["bad"]) == r["good"] + r["bad"], (i, out[i]["good"], out[i]["bad"], r["good"], r["bad"])
        assert abs(int(out[i]["good"]) - r["good"]) <= 2, (i, out[i]["good"], r["good"])
    print(f"fixed-iteration population: worst {worst_r:.2e} rad {worst_t:.2e} m over {N_PAIRS} pairs x {LEVELS} levels x {n_tries} tries")
    with open(os.path.join("gpurun_out", "parity_population_fixed.json"), "w") as f:
        json.dump(dict(n_pairs=N_PAIRS, levels=LEVELS, lm_tries_per_level=n_tries, worst_rot_rad=worst_r, worst_trans_m=worst_t), f)
