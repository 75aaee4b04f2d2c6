#!/usr/bin/env python
"""One short GPU run (numpy + ctypes only, no torch): A/B of the tracking engines on 128 pre-rendered VGA pairs
(scratch/shot_data.npz, 8 distinct streams x 16 replicas in separate device buffers).  For every configuration: median
kernel time (CUDA events inside the library), algorithmic GB/s and whether the results are BIT-identical to the default
cluster engine.  Lines are appended to gpurun_out/shot.jsonl as they are produced.

Engine 4 in this script is scratch/experiments/track_deep.cu (a deeper per-thread gather pipeline).  It measured SLOWER
than the committed cluster engine (profiles/r1_track_engine_ab_shot.jsonl) and is not part of the library; to repeat the
run, add the file to revo_b200/build.py:SOURCES and dispatch engine 4 to launch_track_deep in capi.cu:run_track.

  python scratch/make_shot_data.py 8 && python scratch/final_shot.py
"""
import json
import os
import sys
import time

import numpy as np

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "shot.jsonl"), "a")
BUDGET_S = float(os.environ.get("SHOT_BUDGET_S", "40"))


def emit(**kw):
    kw["t"] = round(time.time() - T0, 2)
    LOG.write(json.dumps(kw) + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())
    print(json.dumps(kw), flush=True)


ENV_KEYS = ("REVO_DEEP_DEPTH", "REVO_DEEP_MINBLOCKS", "REVO_DEEP_SLIM", "REVO_DEEP_HINT", "REVO_TRACK_PCAP", "REVO_TRACK_MAX_CLUSTERS", "REVO_TRACK_PROF")


def main():
    from revo_b200 import api

    z = np.load(os.path.join(ROOT, "scratch", "shot_data.npz"))
    bgr, d16, cam = z["bgr"], z["d16"], z["cam"]
    S = bgr.shape[1]
    reps = int(os.environ.get("SHOT_REPLICAS", "16"))
    B = S * reps
    h, w = bgr.shape[2], bgr.shape[3]
    fx, fy, cx, cy = (float(v) for v in cam[:4])
    st = api.ImgPyramidSettings(PYR_MIN_LVL=3, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    ctx = api.Context(0)
    emit(stage="context", streams=S, pairs=B)
    kf_bgr = np.ascontiguousarray(np.tile(bgr[0], (reps, 1, 1, 1)))
    kf_d = np.ascontiguousarray(np.tile(d16[0], (reps, 1, 1)))
    cu_bgr = np.ascontiguousarray(np.tile(bgr[1], (reps, 1, 1, 1)))
    cu_d = np.ascontiguousarray(np.tile(d16[1], (reps, 1, 1)))
    kf = api.PyramidBatch(ctx, st, kf_bgr, kf_d, B)
    kf.makeKeyframes()
    cur = api.PyramidBatch(ctx, st, cu_bgr, cu_d, B)
    ctx.synchronize()
    tm = ctx.last_timings()
    emit(stage="pyramids", pyr_ms=tm[0], kf_ms=tm[1])
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    Rs = np.tile(np.eye(3, dtype=np.float32), (B, 1, 1))
    Ts = np.zeros((B, 3), np.float32)
    base = {}

    def run(name, engine, C=0, T=0, env=None, n=B, nrep=6, trace_cap=0, key="full"):
        if time.time() - T0 > BUDGET_S:
            emit(name=name, skipped="budget")
            return None
        for k in ENV_KEYS:
            os.environ.pop(k, None)
        os.environ.update(env or {})
        try:
            ctx.set_track_engine(engine, 0)
            ctx.set_track_shape(C, T)
            refs = kf if n == B else [kf[i] for i in range(n)]
            curs = cur if n == B else [cur[i] for i in range(n)]
            ms = []
            out = traces = None
            for r in range(nrep + 2):
                res = trk.trackFramesBatch(Rs[:n], Ts[:n], refs, curs, trace_cap=trace_cap)
                out, traces = res if trace_cap else (res, None)
                if r >= 2:
                    ms.append(ctx.last_timings()[2])
        except Exception as e:  # noqa: BLE001
            emit(name=name, error=str(e)[:300])
            return None
        k_ms = float(np.median(ms))
        pe = int((out["n_evals"].astype(np.int64) * out["n_pts"].astype(np.int64)).sum())
        ev = int(out["n_evals"].sum())
        gbs = 60.0 * pe / (k_ms * 1e-3) / 1e9
        rec = dict(name=name, engine=engine, env=env or {}, n=n, kernel_ms=round(k_ms, 4), min_ms=round(float(min(ms)), 4),
                   evals=ev, gn_iters_per_s=round(ev / (k_ms * 1e-3)), alg_gbs=round(gbs, 1), rc_bad=int((out["rc"] != 0).sum()))
        bkey = (key, trace_cap)
        if bkey not in base:
            base[bkey] = (out.copy(), traces)
            rec["base"] = True
        else:
            b_out, b_tr = base[bkey]
            same = [out[i].tobytes() == b_out[i].tobytes() for i in range(n)]
            rec["bit_identical_pairs"] = int(sum(same))
            rec["bit_identical"] = bool(all(same))
            if not all(same):
                rec["max_dR"] = float(np.abs(out["R"].reshape(n, -1) - b_out["R"].reshape(n, -1)).max())
                rec["max_dt"] = float(np.abs(out["t"] - b_out["t"]).max())
                rec["same_evals"] = int((out["n_evals"] == b_out["n_evals"]).all(axis=1).sum())
            if trace_cap:
                rec["traces_identical"] = bool(traces == b_tr)
        emit(**rec)
        return out

    o = run("cluster C8 T128 (default)", 1)
    if o is not None:
        tot = o["n_evals"].sum(axis=1)
        emit(stage="workload", mean_pts=o["n_pts"].mean(axis=0).tolist(), mean_evals=o["n_evals"].mean(axis=0).tolist(),
             evals_min=int(tot.min()), evals_max=int(tot.max()))
    def deep_env(mb, d, slim=1, hint=0, **kw):
        e = {"REVO_DEEP_MINBLOCKS": str(mb), "REVO_DEEP_DEPTH": str(d), "REVO_DEEP_SLIM": str(slim), "REVO_DEEP_HINT": str(hint)}
        e.update(kw)
        return e

    for mb, d in ((4, 4), (3, 4), (3, 5)):
        run(f"deep mb{mb} d{d}", 4, env=deep_env(mb, d))
    for hint in (1, 2, 3, 4):
        run(f"deep mb4 d4 hint{hint}", 4, env=deep_env(4, 4, hint=hint))
    for mb, d, slim in ((4, 3, 1), (3, 6, 1), (4, 4, 0), (3, 4, 0), (3, 5, 0)):
        run(f"deep mb{mb} d{d} slim{slim}", 4, env=deep_env(mb, d, slim))
    for hint in (1, 2, 3, 4):
        run(f"deep mb3 d5 hint{hint}", 4, env=deep_env(3, 5, hint=hint))
    for maxc in (64, 56, 48):
        run(f"cluster maxc{maxc}", 1, env={"REVO_TRACK_MAX_CLUSTERS": str(maxc)})
        run(f"deep mb4 d4 maxc{maxc}", 4, env={"REVO_TRACK_MAX_CLUSTERS": str(maxc)})
    for maxc in (48, 43):
        run(f"deep mb3 d5 maxc{maxc}", 4, env={"REVO_DEEP_MINBLOCKS": "3", "REVO_DEEP_DEPTH": "5", "REVO_TRACK_MAX_CLUSTERS": str(maxc)})
    # traces (LM decisions) and the small-batch shape
    run("cluster trace", 1, trace_cap=64, nrep=1)
    run("deep mb4 d4 trace", 4, trace_cap=64, nrep=1)
    run("deep mb3 d5 trace", 4, env={"REVO_DEEP_MINBLOCKS": "3", "REVO_DEEP_DEPTH": "5"}, trace_cap=64, nrep=1)
    run("cluster n8 (T256)", 1, n=8, key="n8", nrep=3)
    run("deep n8 T128", 4, n=8, key="n8", nrep=3)
    run("deep n8 T256", 4, C=8, T=256, env={"REVO_DEEP_MINBLOCKS": "2"}, n=8, key="n8", nrep=3)
    run("cluster n1", 1, n=1, key="n1", nrep=3)
    run("deep n1", 4, n=1, key="n1", nrep=3)
    # phase cycle counters (printed by the library on stderr)
    run("cluster prof", 1, env={"REVO_TRACK_PROF": "1"}, nrep=1)
    run("deep mb4 d4 prof", 4, env={"REVO_TRACK_PROF": "1"}, nrep=1)
    run("deep mb3 d5 prof", 4, env={"REVO_TRACK_PROF": "1", "REVO_DEEP_MINBLOCKS": "3", "REVO_DEEP_DEPTH": "5"}, nrep=1)
    emit(stage="done")


if __name__ == "__main__":
    main()
