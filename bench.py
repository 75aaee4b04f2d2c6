#!/usr/bin/env python
"""bench.py -- frames/s and GN-iterations/s of the edge-based RGB-D tracking hot path on B200.

Workload (BASELINE.json configs[1]): TUM-fr1-style synthetic VGA streams, 4-level pyramid, Canny edges.
One STEP = every one of the B independent streams on a GPU advances by one frame: ImgPyramidRGBD
construction (gray, Canny, depth pyramid, 3-D edge lists) for B frames, coarse-to-fine GN/LM tracking of the
B frames against their keyframes in one persistent kernel, and -- every `kf_interval` steps -- keyframe
promotion (exact EDT + lookup structure).  Streams shard over GPUs with no collective (weak scaling).

  value  : frames/s, inputs already resident in HBM when the timed region starts
  e2e    : frames/s through the same public API with pinned HOST buffers (H2D of every frame and D2H of
           every pose inside the timed region)
  roofline: the persistent residual/Jacobian/normal-equation kernel (k_track), algorithmic 60 B per edge
           point per evaluation / its CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline / --impl reference: the CPU oracle (port of the reference's OpenCV/Eigen path; the reference
           itself cannot be built in this image) on the box's host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=256, help="independent streams (frames per step) per GPU")
    ap.add_argument("--ref-streams", type=int, default=64, help="streams per step of the CPU reference arm / cpu_baseline")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--kf-interval", type=int, default=10)
    ap.add_argument("--track-max-clusters", type=int, default=-1,
                    help="resident-cluster cap of the tracking kernel in the pipelined runs (-1 = the library's default for pipelines, 0 = none)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--kf-policy", choices=("interval", "vote"), default="interval",
                    help="interval: a keyframe every --kf-interval frames (the fixed workload of BASELINE.json configs[1]); vote: the "
                         "reference's own policy (assessTrackingQuality after every alignment, promote the previous frame and align "
                         "again on NEW_KF, system.cpp:199-239)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e run without the second (upload/build) stream")
    ap.add_argument("--ctas-per-pair", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true", help="skip the single_stream / config3 / edge_split blocks")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        self.index = index
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU oracle backend (reference arm / cpu_baseline ONLY -- never on the product path)
# ---------------------------------------------------------------------------------------------
def run_cpu(args, bgr_h, depth_h, cam, n_streams, steps, warmup, fidx=lambda i: i):
    """Times `steps` steps of `n_streams` streams on the host cores. bgr_h/depth_h: numpy (F, S, ...)."""
    from oracle.stream_backend import OracleBackend      # reference arm / cpu_baseline ONLY -- never on the product path
    from revo_b200.stream import StreamTracker

    be = OracleBackend(cam, args.levels)
    st = StreamTracker(be, n_streams, args.kf_interval, args.kf_policy)
    st.keep_history = True
    st.start(bgr_h[0, :n_streams], depth_h[0, :n_streams])
    for i in range(1, warmup + 1):
        st.step(bgr_h[fidx(i), :n_streams], depth_h[fidx(i), :n_streams])
    ev0 = st.total_evals
    t0 = time.perf_counter()
    for i in range(warmup + 1, warmup + 1 + steps):
        st.step(bgr_h[fidx(i), :n_streams], depth_h[fidx(i), :n_streams])
    dt = time.perf_counter() - t0
    return dict(seconds=dt, frames=steps * n_streams, evals=st.total_evals - ev0, T_w_c=st.T_w_c.copy(), history=st.history,
                omp_threads=be.omp_threads, cv2_threads=be.cv2_threads)


def pose_errors(T_est, poses_gt, frame):
    """Mean rotation (rad) / translation (m) error of the estimated world poses against the renderer's ground truth
    (both relative to each stream's first frame)."""
    from revo_b200 import synth

    er, et = [], []
    for s in range(T_est.shape[0]):
        T_gt = np.linalg.inv(poses_gt[s][0]) @ poses_gt[s][frame]
        D = np.linalg.inv(T_gt) @ T_est[s].astype(np.float64)
        er.append(synth.rot_angle(D[:3, :3]))
        et.append(float(np.linalg.norm(D[:3, 3])))
    return float(np.mean(er)), float(np.mean(et))


def _render_pair(torch, synth_torch, local_rank, w, h, seed, gap=1):
    dev = torch.device("cuda", local_rank)
    bgr = torch.empty((gap + 1, 1, h, w, 3), dtype=torch.uint8, device=dev)
    depth = torch.empty((gap + 1, 1, h, w), dtype=torch.float32, device=dev)
    cam, poses = synth_torch.render_streams([seed], gap + 1, w, h, dev, bgr, depth)
    return cam, poses, bgr, depth


def extra_single_pair(torch, api, synth_torch, ctx, local_rank, w, h, levels, seed, reps, label):
    """One frame pair at a time (BASELINE configs[0] / configs[2]): latency of a frame = pyramid construction + coarse-to-fine
    tracking against a keyframe, inputs resident in HBM, and the tracking kernel's microseconds per evaluation."""
    cam, poses, bgr, depth = _render_pair(torch, synth_torch, local_rank, w, h, seed)
    fx, fy, cx, cy, _, _ = cam
    st = api.ImgPyramidSettings(PYR_MIN_LVL=levels - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    kf = api.PyramidBatch(ctx, st, bgr[0], depth[0], 1)
    t0 = time.perf_counter()
    kf.makeKeyframes()
    ctx.synchronize()
    kf_first_ms = (time.perf_counter() - t0) * 1e3
    I, Z = np.eye(3, dtype=np.float32)[None], np.zeros((1, 3), np.float32)
    frame_ms, build_ms, k_ms, evals = [], [], [], []
    out = None
    for r in range(reps + 3):
        t0 = time.perf_counter()
        cur = api.PyramidBatch(ctx, st, bgr[1], depth[1], 1)
        out = trk.trackFramesBatch(I, Z, kf, cur)            # synchronous: returns the pose
        dt_ms = (time.perf_counter() - t0) * 1e3
        p_ms, _, t_ms = ctx.last_timings()
        cur.destroy()
        if r >= 3:
            frame_ms.append(dt_ms); build_ms.append(p_ms); k_ms.append(t_ms); evals.append(int(out["n_evals"].sum()))
    t0 = time.perf_counter()
    kf2 = api.PyramidBatch(ctx, st, bgr[1], depth[1], 1)
    kf2.makeKeyframes()
    ctx.synchronize()
    kf2.destroy()
    kf.destroy()
    med = lambda a: float(np.median(a))      # noqa: E731
    ev = med(evals)
    return {"workload": label, "levels": levels, "edge_points_per_level": [int(x) for x in out["n_pts"][0][:levels]],
            "evals_per_pair": ev, "frame_latency_ms": med(frame_ms), "frames_per_sec_single_stream": 1e3 / med(frame_ms),
            "pyramid_build_ms": med(build_ms), "track_kernel_ms": med(k_ms), "us_per_eval": 1e3 * med(k_ms) / ev,
            "gn_iters_per_sec": ev / (med(frame_ms) * 1e-3), "gn_iters_per_sec_tracking_kernel": ev / (med(k_ms) * 1e-3),
            "keyframe_promotion_ms_first_call": kf_first_ms,
            "note": "latency = host wall clock of create + track through the API with a synchronous result (device-resident input); the "
                    "working set of one pair is L2-resident, so these are latency figures, not roofline ones"}


def extra_edge_split(torch, dist, api, synth_torch, ctx, rank, world, local_rank):
    """BASELINE configs[4]: ONE 1920x1080 pair, 3 levels, the edge-point set split over the ranks (rank r evaluates its share of
    every level's list) with one 32-value exchange per evaluation over NVLink inside the persistent kernel."""
    w, h, levels = 1920, 1080, 3
    cam, poses, bgr, depth = _render_pair(torch, synth_torch, local_rank, w, h, 5)
    fx, fy, cx, cy, _, _ = cam
    st = api.ImgPyramidSettings(PYR_MIN_LVL=levels - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
    trk = api.TrackerNew(ctx, api.TrackerSettings(), st)
    kf = api.PyramidBatch(ctx, st, bgr[0], depth[0], 1)
    kf.makeKeyframes()
    cur = api.PyramidBatch(ctx, st, bgr[1], depth[1], 1)
    ctx.synchronize()
    I, Z = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
    # the same pair on one GPU (every rank does it; rank 0 reports)
    single_ms, single_ev = [], 0
    for r in range(6):
        trk.trackFrames(I, Z, kf[0], cur[0])
        if r >= 2:
            single_ms.append(ctx.last_timings()[2])
        single_ev = int(sum(trk.last_result.n_evals))
    R1, T1 = api._R_from_c(np.array(trk.last_result.R)), np.array(trk.last_result.t)
    blobs = [None] * world
    dist.all_gather_object(blobs, trk.splitExport(rank, world))
    trk.splitOpen(b"".join(blobs))
    dist.barrier()
    split_ms, split_ev, rc = [], 0, 0
    R2 = T2 = None
    try:
        for r in range(8):
            dist.barrier()
            _, R2, T2, _ = trk.trackFramesSplit(I, Z, kf[0], cur[0])
            if r >= 3:
                split_ms.append(ctx.last_timings()[2])
            split_ev = int(sum(trk.last_result.n_evals))
    except api.RevoError as e:      # REVO_ERR_COMM: a peer did not answer
        rc = e.code
    t = torch.tensor([float(np.median(split_ms)) if split_ms else 0.0, float(rc)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    poses_all = [None] * world
    dist.all_gather_object(poses_all, None if R2 is None else (np.asarray(R2, np.float32).tobytes(), np.asarray(T2, np.float32).tobytes()))
    kf.destroy()
    cur.destroy()
    if rank != 0:
        return None
    from revo_b200 import synth

    ms, s_ms = float(t[0].item()), float(np.median(single_ms))
    same = all(p is not None and p == poses_all[0] for p in poses_all)
    d_rot = d_trn = None
    if R2 is not None:
        d_rot = synth.rot_angle(np.asarray(R1, np.float64).T @ np.asarray(R2, np.float64))
        d_trn = float(np.linalg.norm(np.asarray(T1, np.float64) - np.asarray(T2, np.float64)))
    return {"workload": "BASELINE.json configs[4]: one 1920x1080 pair, 3 levels, edge-point set split over the ranks, one 32-value "
                        "exchange per evaluation over NVLink inside the persistent kernel (peer-mapped mailboxes, no NCCL call, no host "
                        "round trip)",
            "ranks": world, "edge_points_per_level": [int(x) for x in trk.last_result.n_pts[:levels]], "rc_max_over_ranks": int(t[1].item()),
            "evals_split": split_ev, "evals_single_gpu": single_ev,
            "kernel_ms_split_max_over_ranks": ms, "us_per_eval_split": 1e3 * ms / max(split_ev, 1),
            "kernel_ms_single_gpu": s_ms, "us_per_eval_single_gpu": 1e3 * s_ms / max(single_ev, 1),
            "evals_per_sec_split": split_ev / (ms * 1e-3) if ms > 0 else None,
            "speedup_vs_single_gpu": (s_ms / max(single_ev, 1)) / (ms / max(split_ev, 1)) if ms > 0 else None,
            "bit_identical_across_ranks": bool(same), "pose_vs_single_gpu": {"rot_rad": d_rot, "trans_m": d_trn},
            "limiter": "latency of the per-evaluation exchange: every rank posts 256 B into every peer's mailbox (st.release.sys) and spins "
                       "on its own (ld.acquire.sys); a 1080p level-0 evaluation on one GPU is ~10 us of gather, so the NVLink round "
                       "trip (~2-4 us) bounds the speed-up, not bandwidth"}


def parity_block(gpu_hist, cpu_hist, n_streams):
    """GPU vs float32 CPU oracle on the SAME streams, frames and keyframe schedule: distance of the tracker outputs (pose of the
    frame relative to its keyframe) per stream and step.  Step 1 starts from identical inputs on both sides (identity
    initial guess); later steps chain each side's own previous result through the motion model."""
    from revo_b200 import synth

    n = min(len(gpu_hist), len(cpu_hist))
    rot, trn, same = [], [], []
    for k in range(n):
        Tg, Tc = gpu_hist[k][1][:n_streams].astype(np.float64), cpu_hist[k][1][:n_streams].astype(np.float64)
        eg, ec = gpu_hist[k][2][:n_streams], cpu_hist[k][2][:n_streams]
        for s_ in range(n_streams):
            D = np.linalg.inv(Tc[s_]) @ Tg[s_]
            rot.append(synth.rot_angle(D[:3, :3]))
            trn.append(float(np.linalg.norm(D[:3, 3])))
            same.append(bool((eg[s_] == ec[s_]).all()))
    rot, trn, same = np.array(rot), np.array(trn), np.array(same)
    first = slice(0, n_streams)
    q = lambda a, p_: float(np.percentile(a, p_))      # noqa: E731
    out = {"n_streams": n_streams, "n_frames": n, "reference": "float32 CPU oracle (reference-as-is arithmetic), same frames and keyframe schedule",
           "rot_rad": {"median": q(rot, 50), "p95": q(rot, 95), "max": float(rot.max())},
           "trans_m": {"median": q(trn, 50), "p95": q(trn, 95), "max": float(trn.max())},
           "same_evals_fraction": float(same.mean()),
           "first_frame": {"rot_rad_median": q(rot[first], 50), "rot_rad_max": float(rot[first].max()), "trans_m_median": q(trn[first], 50),
                           "trans_m_max": float(trn[first].max()), "same_evals_fraction": float(same[first].mean())},
           "note": "default termination rules: the accept / convergence tests of the LM loop sit on float-rounding knife edges, so the "
                   "evaluation counts agree on a fraction of the pairs only (the float32 and float64 oracles differ from each other the "
                   "same way, tests/test_gpu_parity_population.py); where they agree the poses match to ~1e-6"}
    if same.any():
        out["same_evals"] = {"rot_rad_max": float(rot[same].max()), "trans_m_max": float(trn[same].max())}
    return out


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and therefore its pinned host buffers, allocated first-touch afterwards) to the NUMA node the rank's GPU
    hangs off, so that the host->device copies of the end-to-end run do not cross the socket interconnect.  Returns a dict for
    the JSON line (node, cpus) or the reason it did nothing."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip().lower()
        dom, rest = out.split(":", 1)
        dev = f"{dom[-4:]}:{rest}"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return {"bound": False, "why": "numa_node = -1 (single node or not reported)"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": f"no allowed cpu on node {node}"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus), "gpu_pci": dev}
    except Exception as e:      # no nvidia-smi / sysfs entry: run unbound
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


def main():
    args = parse_args()
    numa = bind_to_gpu_numa_node(int(os.environ.get("LOCAL_RANK", "0"))) if args.impl != "reference" and not args.no_numa else {"bound": False, "why": "disabled"}
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, K = max(args.warmup, 0), max(args.steps, 1)
    # frames rendered per stream; longer runs walk the stream back and forth (camera reverses), which bounds host memory
    n_frames = min(1 + W + K, 16)

    def fidx(i):
        """frame index of step i on the ping-pong path 0,1,..,n-1,n-2,..,1,0,1,.."""
        period = 2 * (n_frames - 1)
        j = i % period
        return j if j < n_frames else period - j
    w, h = args.width, args.height

    # ------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        from revo_b200 import synth_torch

        S = args.ref_streams
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        bgr = torch.empty((n_frames, S, h, w, 3), dtype=torch.uint8)
        depth = torch.empty((n_frames, S, h, w), dtype=torch.float32)
        cam, poses = synth_torch.render_streams([2000 + s for s in range(S)], n_frames, w, h, dev, bgr, depth)
        from oracle import oracle as O

        O.build()
        import cv2

        depth16 = np.round(depth.numpy().astype(np.float64) * 5000.0).astype(np.uint16)   # the wire format (TUM 16-bit depth)
        r = run_cpu(args, bgr.numpy(), depth16, cam, S, K, W, fidx)
        er, et = pose_errors(r["T_w_c"], poses, fidx(W + K))
        fps = r["frames"] / r["seconds"]
        cores = os.cpu_count()
        line = {
            "impl": "reference", "metric": "frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * r["seconds"] / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gn_iters_per_sec": r["evals"] / r["seconds"],
            "config": {"workload": f"TUM-fr1-style synthetic {w}x{h} RGB-D streams, {args.levels}-level pyramid, Canny 150/100, "
                                   f"keyframe every {args.kf_interval} frames; bounded sample: {S} streams x {K} frames per run",
                       "streams_per_step": S},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "omp_threads": r["omp_threads"], "cv2_threads": r["cv2_threads"],
                             "sample": f"{S} streams x {K} timed frames ({r['frames']} frames); OpenCV kernels via cv2 "
                                       f"{cv2.__version__} ({r['cv2_threads']} threads), loops + tracker = C port of the "
                                       f"reference (OpenMP over independent pairs, {r['omp_threads']} threads, set explicitly)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "pose_error_vs_ground_truth": {"rot_rad": er, "trans_m": et},
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------- our arm
    from revo_b200 import api, synth_torch
    from revo_b200.stream import CudaBackend, StreamTracker

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.streams
    ctx = api.Context(local_rank)
    if args.ctas_per_pair or args.threads:
        ctx.set_track_shape(args.ctas_per_pair, args.threads)
    ext_stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # ---- synthetic inputs: B distinct streams per rank, rendered on the GPU into pinned host memory ----
    bgr_h = torch.empty((n_frames, B, h, w, 3), dtype=torch.uint8).pin_memory()
    depth_h = torch.empty((n_frames, B, h, w), dtype=torch.float32).pin_memory()
    seeds = [2000 + rank * B + s for s in range(B)]
    cam, poses = synth_torch.render_streams(seeds, n_frames, w, h, torch.device("cuda", local_rank), bgr_h, depth_h)
    bgr_d = bgr_h.to(torch.device("cuda", local_rank))
    depth_d = depth_h.to(torch.device("cuda", local_rank))
    # the sensor / dataset wire format of the depth (TUM: 16-bit, 5000 units per metre): what the end-to-end run uploads
    depth16_h = torch.empty((n_frames, B, h, w), dtype=torch.int16).pin_memory()
    for fi in range(n_frames):
        z = (depth_d[fi].double() * 5000.0).round().to(torch.int32)
        back = (z.double() * float(np.float32(1.0) / np.float32(5000.0))).float()
        assert torch.equal(back, depth_d[fi]), "16-bit depth does not reproduce the float depth bit for bit"
        depth16_h[fi].copy_(z.to(torch.int16))     # two's-complement wrap: the same 16 bits as uint16
    torch.cuda.synchronize()

    fx, fy, cx, cy, _, _ = cam
    settings = api.ImgPyramidSettings(PYR_MIN_LVL=args.levels - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    build_ctx = api.Context(local_rank)    # second stream: upload + pyramid build of frame k+1 overlap tracking of frame k

    def timed_run(src_bgr, src_depth, sample_clocks, pipelined=False, K=K, keep_history=False, policy=None):
        be = CudaBackend(ctx, settings, build_ctx=build_ctx if pipelined else None,
                         track_max_clusters=None if args.track_max_clusters < 0 else args.track_max_clusters)
        st = StreamTracker(be, B, args.kf_interval, policy or args.kf_policy)
        st.keep_history = keep_history
        # rank 0 samples its GPU every 20 ms; the other ranks every 250 ms (eight 50 Hz nvidia-smi loops on one box disturb the
        # ranks' host threads) -- their medians and throttle reasons are merged into the reported block below
        sampler = ClockSampler(local_rank, 20 if rank == 0 else 250) if sample_clocks else None
        if sampler:
            sampler.start()     # sampled from the warm-up on: the GPU is under the same load throughout
        st.start(src_bgr[0], src_depth[0])
        if pipelined:   # two frames in flight: upload of k+2 | build of k+1 | track of k
            st.prefetch(src_bgr[fidx(1)], src_depth[fidx(1)])
            st.prefetch(src_bgr[fidx(2)], src_depth[fidx(2)])
        for i in range(1, W + 1):
            if pipelined:
                st.step_pipelined(src_bgr[fidx(i + 2)], src_depth[fidx(i + 2)])
            else:
                st.step(src_bgr[fidx(i)], src_depth[fidx(i)])
        ctx.synchronize()
        build_ctx.synchronize()
        ev0, pe0, l0 = st.total_evals, st.total_point_evals, ctx.launch_count + build_ctx.launch_count
        k9_ms = pyr_ms = kf_ms = 0.0
        step_wall = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext_stream)
        for i in range(W + 1, W + 1 + K):
            ts = time.perf_counter()
            if pipelined:
                st.step_pipelined(src_bgr[fidx(i + 2)], src_depth[fidx(i + 2)])   # the prefetches beyond the last step are part of the cost
            else:
                st.step(src_bgr[fidx(i)], src_depth[fidx(i)])
            if not pipelined:     # per-phase CUDA-event times: only in the serial pass (querying them synchronises the streams)
                p, kf, k9 = ctx.last_timings()
                pyr_ms += p
                k9_ms += k9
                if st.kf_policy == "interval" and st.frame % args.kf_interval == 0:
                    kf_ms += kf
            step_wall.append((time.perf_counter() - ts) * 1e3)
        build_ctx.synchronize()     # the uploads / builds enqueued by the last steps are part of the timed region
        e1.record(ext_stream)
        upload_ms = build_ctx.last_upload_ms() if pipelined else ctx.last_upload_ms()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if clocks is not None and dist is not None:
            every = [None] * world
            dist.all_gather_object(every, clocks)
            meds = [c["sm_mhz"] for c in every if c and c.get("sm_mhz") is not None]
            clocks = dict(clocks, sm_mhz=min(meds) if meds else clocks.get("sm_mhz"),
                          reasons=sorted(set(r for c in every if c for r in c.get("reasons", []))),
                          per_rank_sm_mhz=[c.get("sm_mhz") if c else None for c in every])
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res = dict(ms=ms, wall=wall, evals=st.total_evals - ev0, point_evals=st.total_point_evals - pe0,
                   launches=ctx.launch_count + build_ctx.launch_count - l0, step_wall=step_wall, upload_ms=upload_ms, k9_ms=k9_ms, pyr_ms=pyr_ms, kf_ms=kf_ms, T_w_c=st.T_w_c.copy(), clocks=clocks,
                   n_pts=st.last["n_pts"].mean(axis=0).tolist(), n_evals=st.last["n_evals"].mean(axis=0).tolist(), history=st.history,
                   n_keyframes=st.n_keyframes, n_retracks=st.n_retracks)
        st.close()
        return res

    def note(msg):
        if rank == 0:
            sys.stderr.write(f"[bench] {time.strftime('%H:%M:%S')} {msg}\n")
            sys.stderr.flush()

    note(f"inputs rendered: {n_frames} frames x {B} streams")
    # inputs resident in HBM; the same two-stream pipeline as the end-to-end run (build of frame k+1 overlaps tracking of k)
    dev_run = timed_run(bgr_d, depth_d, sample_clocks=True, pipelined=not args.no_pipeline, keep_history=True)
    note("device-resident run done; host ms per step: " + " ".join(f"{x:.2f}" for x in dev_run["step_wall"]))
    # pinned host inputs, H2D inside the timed region; the upload + pyramid build of frame k+1 run on a second stream
    # while frame k is tracked (same public API, two contexts)
    host_run = timed_run(bgr_h, depth16_h, sample_clocks=False, pipelined=not args.no_pipeline)
    host_run_f32 = timed_run(bgr_h, depth_h, sample_clocks=False, pipelined=not args.no_pipeline)   # float-depth interface
    note("host-input (e2e) run done; host ms per step: " + " ".join(f"{x:.2f}" for x in host_run["step_wall"]))
    # kernel-level numbers (phase times, roofline of k_track) from a pass in which nothing else runs beside the kernel being
    # timed: same workload, device-resident, one stream, no overlap
    iso_run = timed_run(bgr_d, depth_d, sample_clocks=False, pipelined=False, K=min(K, 10))
    K_iso = min(K, 10)
    note("isolated-kernel pass done")

    # what the host link can do: one pinned -> device copy of a batch of bgr frames, timed alone
    # host link: every rank copies pinned frames at the same time (behind a barrier), the way the end-to-end run loads the box
    tmp = torch.empty_like(bgr_d[0])
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tmp.copy_(bgr_h[0], non_blocking=True)
    barrier()
    c0.record()
    for r in range(4):
        tmp.copy_(bgr_h[1 + r % (n_frames - 1)], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1)
    if dist is not None:
        t = torch.tensor([h2d_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_ms = float(t.item())
    h2d_gbs = 4 * bgr_h[1].numel() / (h2d_ms * 1e-3) / 1e9        # per rank, all ranks copying concurrently
    del tmp

    frames_rank = K * B
    tot = torch.tensor([float(dev_run["evals"]), float(dev_run["point_evals"]), float(dev_run["launches"])], device="cuda",
                       dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot)
    evals_all, point_evals_all, launches_all = [float(x) for x in tot.tolist()]
    frames_all = frames_rank * world
    value = frames_all / (dev_run["ms"] * 1e-3)
    e2e = frames_all / (host_run["ms"] * 1e-3)
    er, et = pose_errors(dev_run["T_w_c"], poses, fidx(W + K))

    # roofline of the dominant kernel (rank 0's launches): algorithmic 60 B / point / evaluation
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    k9_s = iso_run["k9_ms"] * 1e-3
    achieved = 60.0 * iso_run["point_evals"] / k9_s / 1e9 if k9_s > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_track_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    edge_split = None
    if world > 1 and not args.no_extras:
        edge_split = extra_edge_split(torch, dist, api, synth_torch, ctx, rank, world, local_rank)
        note("edge-split block done")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    line = {
        "metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_run["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"TUM-fr1-style synthetic {w}x{h} RGB-D streams, {args.levels}-level pyramid, Canny 150/100, "
                               f"keyframe every {args.kf_interval} frames (BASELINE.json configs[1])",
                   "streams_per_gpu": B, "frames_per_step": B * world, "parallelism": f"stream-shard x{world}, no collective",
                   "l2_policy": f"inputs larger than L2: every step touches {B} new frames "
                                f"({B * w * h * 7 / 1e6:.0f} MB of bgr+depth) plus {B} keyframe structures"},
        "gn_iters_per_sec": evals_all / (dev_run["ms"] * 1e-3),
        "gn_iters_per_sec_tracking_kernel": iso_run["evals"] / k9_s if k9_s > 0 else None,
        "gpu_launches": int(launches_all),
        "phase_ms_per_step": {"pyramid": iso_run["pyr_ms"] / K_iso, "keyframe": iso_run["kf_ms"] / K_iso,
                              "track_kernel": iso_run["k9_ms"] / K_iso, "whole_step_serial": iso_run["ms"] / K_iso,
                              "whole_step_pipelined": dev_run["ms"] / K,
                              "note": "phases timed with CUDA events in a separate pass without stream overlap; the value run "
                                      "overlaps the pyramid build of frame k+1 with the tracking of frame k on two streams"},
        "mean_edge_points_per_level": dev_run["n_pts"], "mean_evals_per_level": dev_run["n_evals"],
        "pose_error_vs_ground_truth": {"rot_rad": er, "trans_m": et},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(B * w * h * 5), "d2h_bytes_per_step": int(B * 128),
                "input": "pinned host bgr u8 + raw 16-bit depth (TUM wire format, converted on the device like the reference's reader "
                         "does on the host); float-depth interface (7 B/px): see float_depth",
                "float_depth": {"value": frames_all / (host_run_f32["ms"] * 1e-3), "unit": "frames/s",
                                "h2d_bytes_per_step": int(B * w * h * 7), "ms_per_step": host_run_f32["ms"] / K},
                "ms_per_step": host_run["ms"] / K, "h2d_link_gbs_measured": h2d_gbs, "upload_ms_last_batch": host_run["upload_ms"],
                "h2d_floor_ms_per_step": B * w * h * 5 / (h2d_gbs * 1e9) * 1e3,
                "h2d_aggregate_gbs_all_ranks": h2d_gbs * world, "h2d_achieved_gbs_all_ranks": world * B * w * h * 5 / (host_run["ms"] / K * 1e-3) / 1e9,
                "host_link_note": "h2d_link_gbs_measured = pinned-host -> device copy rate per rank with ALL ranks copying at once "
                                  "(slowest rank); the end-to-end step cannot be shorter than h2d_floor_ms_per_step, whatever the kernels do",
                "numa": numa},
        "roofline": {"kernel": "k_track (persistent residual/Jacobian/6x6 reduce + LM)", "bound": "hbm", "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "algorithmic_bytes": 60.0 * iso_run["point_evals"] / K_iso,
                     "timing": "CUDA events around the kernel on its launch stream, pass without concurrent kernels",
                     "note": "algorithmic bytes = 60 B x sum(n_pts x n_evals); keyframe structures of a pair stay L2-resident "
                             "across its LM iterations, so achieved may exceed DRAM traffic"},
        "clocks": dev_run["clocks"],
    }

    # pyramid build against ITS roofline: SURVEY 8(d) algorithmic bytes per frame / CUDA-event time of the build phase
    lv_px = [(w >> l) * (h >> l) for l in range(args.levels)]
    pts_per_frame = float(sum(dev_run["n_pts"][:args.levels]))
    pyr_bytes_frame = 16.0 * lv_px[0] + 12.0 * sum(lv_px[1:]) + 16.0 * pts_per_frame
    kf_bytes_frame = 25.0 * sum(lv_px)
    pyr_s, kf_s = iso_run["pyr_ms"] / K_iso * 1e-3, iso_run["kf_ms"] / max(1, sum(1 for i in range(W + 1, W + 1 + K_iso) if i % args.kf_interval == 0)) * 1e-3
    line["roofline_pyramid"] = {
        "bound": "hbm", "unit": "GB/s", "peak": peak,
        "build": {"algorithmic_bytes_per_frame": pyr_bytes_frame, "achieved": pyr_bytes_frame * B / pyr_s / 1e9 if pyr_s > 0 else None,
                  "frac": pyr_bytes_frame * B / pyr_s / 1e9 / peak if pyr_s > 0 else None, "ms_per_batch": pyr_s * 1e3},
        "keyframe": {"algorithmic_bytes_per_frame": kf_bytes_frame, "achieved": kf_bytes_frame * B / kf_s / 1e9 if kf_s > 0 else None,
                     "frac": kf_bytes_frame * B / kf_s / 1e9 / peak if kf_s > 0 else None, "ms_per_promotion": kf_s * 1e3},
        "note": "bytes: 16 B/px level 0, 12 B/px levels >= 1, 16 B per edge point; keyframe promotion 25 B/px (SURVEY.md 8d)"}

    if edge_split is not None:
        line["edge_split"] = edge_split

    if world == 1 and not args.no_extras and args.kf_policy != "vote":
        # the reference's own keyframe policy (vote after every alignment, promote the previous frame + align again on NEW_KF)
        # on the same streams, device-resident inputs, same pipeline: batched votes / promotions / re-alignments
        Kv = min(K, 20)
        vr = timed_run(bgr_d, depth_d, sample_clocks=False, pipelined=not args.no_pipeline, K=Kv, policy="vote")
        line["kf_policy_vote"] = {
            "value": B * Kv / (vr["ms"] * 1e-3), "unit": "frames/s", "ms_per_step": vr["ms"] / Kv, "steps": Kv,
            "gn_iters_per_sec": vr["evals"] / (vr["ms"] * 1e-3),
            "keyframe_switches_per_frame": vr["n_keyframes"] / float(B * (Kv + W)),
            "note": "on these smooth synthetic streams the vote rarely (often never) asks for a new keyframe within the run, so this "
                    "number carries the cost of voting, not of promotions; tests/test_gpu_track.py::test_stream_tracker_vote_policy_on_gpu "
                    "drives streams that do switch and checks decisions and poses against per-stream main loops",
            "policy": "assessTrackingQuality for all streams in one launch pair (revo_track_quality_batch); streams voting NEW_KF "
                      "promote their previous frame, are aligned again and vote again in a second, smaller launch of each kind "
                      "(system.cpp:199-239); one device->host read of the counters per vote"}
        note("vote-policy run done")

    if world == 1 and not args.no_extras:
        line["single_stream"] = extra_single_pair(torch, api, synth_torch, ctx, local_rank, 640, 480, 3, seed=1, reps=30,
                                                  label="BASELINE.json configs[0]: one 640x480 pair stream, 3-level pyramid")
        line["config3"] = extra_single_pair(torch, api, synth_torch, ctx, local_rank, 1280, 960, 5, seed=3, reps=10,
                                            label="BASELINE.json configs[2]: one 1280x960 pair, 5-level pyramid, Huber weights")
        note("single-pair extras done")

    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O

        O.build()
        import cv2

        S = min(args.ref_streams, B)
        r = run_cpu(args, bgr_h.numpy(), depth16_h.numpy().view(np.uint16), cam, S, min(K, 40), 1, fidx)
        line["cpu_baseline"] = {"value": r["frames"] / r["seconds"], "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "omp_threads": r["omp_threads"], "cv2_threads": r["cv2_threads"],
                                "gn_iters_per_sec": r["evals"] / r["seconds"],
                                "sample": f"{S} of the {B} streams x {min(K, 40)} frames ({r['frames']} frames, {r['seconds']:.1f} s); "
                                          f"cv2 {cv2.__version__} ({r['cv2_threads']} threads) + C port of the reference loops/tracker "
                                          f"(OpenMP over pairs, {r['omp_threads']} threads, set explicitly)"}
        # parity on the benchmarked workload: the GPU's poses against the CPU oracle's on the shared streams
        line["parity"] = parity_block(dev_run["history"], r["history"], S)
    print(json.dumps(line))
    sys.stderr.write(f"[bench] {value:.0f} frames/s (e2e {e2e:.0f}), {line['gn_iters_per_sec']:.0f} GN-iters/s, step {dev_run['ms'] / K:.2f} ms = "
                     f"pyr {iso_run['pyr_ms'] / K_iso:.2f} + kf {iso_run['kf_ms'] / K_iso:.2f} + track {iso_run['k9_ms'] / K_iso:.2f} ms (serial {iso_run['ms'] / K_iso:.2f}), "
                     f"k_track {achieved:.0f} GB/s algorithmic = {achieved / peak:.3f} of HBM peak\n")
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
