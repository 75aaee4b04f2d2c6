#!/usr/bin/env python
"""Micro-benchmark of the tracking kernel alone (both engines, several launch shapes) on one batch of synthetic
VGA pairs: kernel ms (CUDA events inside the library), GN-iterations/s, algorithmic GB/s (60 B x points x evals).
Used to pick defaults; bench.py remains the contract benchmark.

  python scratch/track_bench.py --streams 128 --gap 2
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=128)
    ap.add_argument("--gap", type=int, default=1, help="frames between keyframe and tracked frame")
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--configs", default="")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--quick", action="store_true", help="only the selected configs (for ncu)")
    args = ap.parse_args()
    import torch

    from revo_b200 import api, synth_torch

    B, w, h = args.streams, args.width, args.height
    dev = torch.device("cuda", 0)
    nf = args.gap + 1
    bgr = torch.empty((nf, B, h, w, 3), dtype=torch.uint8, device=dev)
    depth = torch.empty((nf, B, h, w), dtype=torch.float32, device=dev)
    cam, poses = synth_torch.render_streams([2000 + s for s in range(B)], nf, w, h, dev, bgr, depth)
    fx, fy, cx, cy, _, _ = cam
    st = api.ImgPyramidSettings(PYR_MIN_LVL=args.levels - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    ctx = api.Context(0)
    kf = api.PyramidBatch(ctx, st, bgr[0], depth[0], B)
    kf.makeKeyframes()
    cur = api.PyramidBatch(ctx, st, bgr[args.gap], depth[args.gap], B)
    ctx.synchronize()
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    Rs = np.tile(np.eye(3, dtype=np.float32), (B, 1, 1))
    Ts = np.zeros((B, 3), np.float32)

    # (name, engine, ctas_per_pair, threads, chunk_points, env)
    configs = [
        ("C8 T128 (default)", 1, 8, 128, 0, {}),
        ("C8 T128 no speculation", 1, 8, 128, 0, {"REVO_TRACK_SPEC": "0"}),
        ("C8 T128 plain gathers", 1, 8, 128, 0, {"REVO_TRACK_HINT": "0"}),
        ("C8 T128 L1::no_allocate", 1, 8, 128, 0, {"REVO_TRACK_HINT": "1"}),
        ("C4 T256", 1, 4, 256, 0, {}),
        ("C8 T256", 1, 8, 256, 0, {}),
        ("C8 T128 pcap8", 1, 8, 128, 0, {"REVO_TRACK_PCAP": "8"}),
        ("C8 T128 pcap18", 1, 8, 128, 0, {"REVO_TRACK_PCAP": "18"}),
        ("C8 T128 maxc64", 1, 8, 128, 0, {"REVO_TRACK_MAX_CLUSTERS": "64"}),
        ("C4 T128", 1, 4, 128, 0, {}),
        ("C4 T128 plain", 1, 4, 128, 0, {"REVO_TRACK_HINT": "0"}),
        ("C4 T128 maxc100", 1, 4, 128, 0, {"REVO_TRACK_MAX_CLUSTERS": "100"}),
        ("C4 T128 maxc120", 1, 4, 128, 0, {"REVO_TRACK_MAX_CLUSTERS": "120"}),
        ("C2 T256", 1, 2, 256, 0, {}),
        ("C4 T128 pcap40", 1, 4, 128, 0, {"REVO_TRACK_PCAP": "40"}),
    ]
    if args.configs:
        keep = set(int(x) for x in args.configs.split(","))
        configs = [c for i, c in enumerate(configs) if i in keep]
    base = None
    for name, eng, C, T, chunk, env in configs:
        for k in ("REVO_Q_OVERSUB_X4", "REVO_Q_SMIN", "REVO_Q_THREADS", "REVO_TRACK_PCAP", "REVO_TRACK_MAX_CLUSTERS", "REVO_TRACK_SPEC", "REVO_TRACK_HINT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            ctx.set_track_shape(C, T)
            ms = []
            for r in range(args.reps + 2):
                t0 = time.perf_counter()
                out = trk.trackFramesBatch(Rs, Ts, kf, cur)
                wall = (time.perf_counter() - t0) * 1e3
                if r >= 2:
                    ms.append((ctx.last_timings()[2], wall))
        except api.RevoError as e:
            print(f"{name:24s} FAILED: {e}")
            continue
        k_ms = float(np.median([m[0] for m in ms]))
        wall_ms = float(np.median([m[1] for m in ms]))
        pe = int((out["n_evals"].astype(np.int64) * out["n_pts"].astype(np.int64)).sum())
        ev = int(out["n_evals"].sum())
        gbs = 60.0 * pe / (k_ms * 1e-3) / 1e9
        Rm = out["R"].reshape(B, 9)
        if base is None:
            base = (Rm.copy(), out["t"].copy(), out["n_evals"].copy())
        dR = float(np.abs(Rm - base[0]).max())
        dT = float(np.abs(out["t"] - base[1]).max())
        same = int((out["n_evals"] == base[2]).all(axis=1).sum())
        print(f"{name:24s} kernel {k_ms:7.3f} ms  wall {wall_ms:7.3f} ms  evals {ev:6d}  {ev / (k_ms * 1e-3) / 1e6:6.2f} M it/s  "
              f"{gbs:7.0f} GB/s alg  ({gbs / 6547.8:.3f} of HBM)  max|dR| {dR:.1e} max|dT| {dT:.1e} same-evals {same}/{B} "
              f"rc!=0 {int((out['rc'] != 0).sum())}", flush=True)
    if args.quick:
        return
    tot = out["n_evals"].sum(axis=1)
    print("evals per pair: min %d p50 %d p90 %d max %d" % (tot.min(), np.median(tot), np.percentile(tot, 90), tot.max()))
    # profile pass (phase cycle counters; slows the kernel slightly) + sub-batches (critical path vs throughput)
    os.environ["REVO_TRACK_PROF"] = "1"
    for k in ("REVO_Q_OVERSUB_X4", "REVO_Q_SMIN", "REVO_Q_THREADS", "REVO_TRACK_PCAP", "REVO_TRACK_MAX_CLUSTERS", "REVO_TRACK_SPEC", "REVO_TRACK_HINT"):
        os.environ.pop(k, None)
    ctx.set_track_shape(0, 0)
    trk.trackFramesBatch(Rs, Ts, kf, cur)
    os.environ.pop("REVO_TRACK_PROF")
    for nb in (1, 8, 32, 64, B):
        sub_k = api.PyramidBatch.__new__(api.PyramidBatch)
        for eng in (1,):
            refs = [kf[i] for i in range(nb)]
            curs = [cur[i] for i in range(nb)]
            ms = []
            for r in range(5):
                o = trk.trackFramesBatch(Rs[:nb], Ts[:nb], refs, curs)
                if r >= 2:
                    ms.append(ctx.last_timings()[2])
            ev = o["n_evals"].sum(axis=1)
            print(f"  n_pairs {nb:4d} engine {eng}: kernel {np.median(ms):7.3f} ms, max evals/pair {ev.max()}, "
                  f"us per eval of the longest pair {1e3 * np.median(ms) / ev.max():.2f}")
    print(json.dumps({"mean_pts": out["n_pts"].mean(axis=0).tolist(), "mean_evals": out["n_evals"].mean(axis=0).tolist()}))


if __name__ == "__main__":
    main()
