#!/usr/bin/env python
"""CPU study for the round-2 lookup-structure layout of k_track (no GPU needed).

Part 1 -- L2 LINE footprint.  k_track gathers one record per edge point and evaluation; a 32-byte record allocates a
128-byte L2 line, and the measured limiter of more pairs in flight is the number of distinct lines a pair touches
(profiles/r1_k_track_v6_hotspots.txt).  For synthetic VGA pairs tracked by the CPU oracle, the points of every level are
projected at the converged pose and the distinct 128-byte lines / 32-byte sectors are counted for candidate layouts.

Part 2 -- what precision the lookup structure needs.  The oracle tracks the same pairs with the {gx, gy, dt} structure
quantised the way a layout would store it; reported: pose difference to the unquantised run, evaluation counts.

  python scratch/layout_study.py [n_pairs]      -> profiles/r1_lookup_layout_study.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from revo_b200 import synth  # noqa: E402

N_LEVELS = 4


def rot_angle(R):
    """Small-angle safe: |vee(R - R^T)| / 2 in float64 (arccos of a float32 trace resolves only ~5e-4 rad)."""
    R = np.asarray(R, np.float64)
    return float(0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]))


# ---- part 1: layouts --------------------------------------------------------------------------------------------
def lines_quad(ix, iy, w, rec_bytes, bw, bh):
    """Quad records (one record per pixel holds its 2x2 neighbourhood), rec_bytes each, stored in blocks of bw x bh pixels
    (bw*bh*rec_bytes == 128 -> one block per line).  Returns (line ids, sector ids)."""
    per_line = 128 // rec_bytes
    assert bw * bh == per_line
    bpr = (w + bw - 1) // bw
    line = (iy // bh) * bpr + ix // bw
    within = (iy % bh) * bw + ix % bw
    sector = line * 4 + (within * rec_bytes) // 32
    return line, sector


def lines_texel(ix, iy, w, tex_bytes, bw, bh):
    """Unique texels (tex_bytes each) in tiles of bw x bh texels per 128-byte line; a point touches its 2x2 neighbourhood."""
    per_line = 128 // tex_bytes
    assert bw * bh == per_line
    bpr = (w + bw - 1) // bw
    ls, ss = [], []
    for dy in (0, 1):
        for dx in (0, 1):
            x, y = ix + dx, iy + dy
            line = (y // bh) * bpr + x // bw
            within = (y % bh) * bw + x % bw
            ls.append(line)
            ss.append(line * 4 + (within * tex_bytes) // 32)
    return np.concatenate(ls), np.concatenate(ss)


LAYOUTS = [
    # name, kind, bytes, block w, block h, loads per point
    ("quad 32 B, row-major 4x1 (round 1)", "quad", 32, 4, 1, "1 x 256 bit"),
    ("quad 32 B, 2x2 blocks", "quad", 32, 2, 2, "1 x 256 bit"),
    ("quad 16 B (u16 dt, s8 grad), 8x1", "quad", 16, 8, 1, "1 x 128 bit"),
    ("quad 16 B, 4x2 blocks", "quad", 16, 4, 2, "1 x 128 bit"),
    ("quad 16 B, 2x4 blocks", "quad", 16, 2, 4, "1 x 128 bit"),
    ("texel 8 B (f32 dt, s16 grad), row-major 16x1", "texel", 8, 16, 1, "4 x 64 bit"),
    ("texel 8 B, 4x4 tiles", "texel", 8, 4, 4, "4 x 64 bit"),
    ("texel 8 B, 8x2 tiles", "texel", 8, 8, 2, "4 x 64 bit"),
    ("texel 4 B (u16 dt, s8 grad), 8x4 tiles", "texel", 4, 8, 4, "4 x 32 bit"),
    ("texel 4 B, row-major 32x1", "texel", 4, 32, 1, "2 x 64 bit"),
]


def project(pts4, cam, R, T):
    X = pts4[:, :3].astype(np.float64) @ np.asarray(R, np.float64).T + np.asarray(T, np.float64)
    u = cam.fx * X[:, 0] / X[:, 2] + cam.cx
    v = cam.fy * X[:, 1] / X[:, 2] + cam.cy
    ok = (u > 1) & (v > 1) & (u < cam.w - 2) & (v < cam.h - 2)
    return u[ok].astype(np.int64), v[ok].astype(np.int64)


# ---- part 2: quantisation ------------------------------------------------------------------------------------------
def q_snorm(g, bits):
    s = float((1 << (bits - 1)) - 4) if bits >= 10 else float((1 << (bits - 1)) - 1)
    return np.rint(np.clip(g, -1, 1) * s) / s


def quantise(opt, mode, ed):
    o = opt.copy()
    gq, dq = mode
    if gq[0] == "s":
        o[..., 0], o[..., 1] = q_snorm(o[..., 0], int(gq[1:])), q_snorm(o[..., 1], int(gq[1:]))
    elif gq == "f16":
        o[..., 0], o[..., 1] = o[..., 0].astype(np.float16), o[..., 1].astype(np.float16)
    if dq == "u16":      # clamp at ed + 2 (safe for the edge filter: the EDT is 1-Lipschitz), 16-bit fixed point
        dc = ed + 2.0
        o[..., 2] = np.rint(np.minimum(o[..., 2], dc) * (65535.0 / dc)) * (dc / 65535.0)
    elif dq == "f16":
        o[..., 2] = o[..., 2].astype(np.float16)
    elif dq == "u8":
        dc = ed + 2.0
        o[..., 2] = np.rint(np.minimum(o[..., 2], dc) * (255.0 / dc)) * (dc / 255.0)
    return o.astype(np.float32)


QUANT = [
    ("s16 grad, f32 dt (round 1 device layout)", ("s16", "f32")),
    ("s16 grad, u16 dt clamped at ed+2", ("s16", "u16")),
    ("s16 grad, f16 dt", ("s16", "f16")),
    ("s12 grad, f32 dt", ("s12", "f32")),
    ("s10 grad, f32 dt", ("s10", "f32")),
    ("s8 grad, f32 dt", ("s8", "f32")),
    ("s8 grad, f16 dt", ("s8", "f16")),
    ("s8 grad, u16 dt clamped at ed+2", ("s8", "u16")),
    ("f16 grad, f16 dt", ("f16", "f16")),
    ("s8 grad, u8 dt (too coarse, for scale)", ("s8", "u8")),
]


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    orc = O.Oracle("f32")
    cfg = orc.default_cfg()
    pcfg = O.PyrCfg(n_levels=N_LEVELS)
    out = []
    P = out.append
    fp = {name: np.zeros((N_LEVELS, 3)) for name, *_ in LAYOUTS}     # lines, sectors, points per level
    qres = {name: [] for name, _ in QUANT}
    base_evals = []
    for s in range(n_pairs):
        st = synth.make_stream(2000 + s, 3, 640, 480)
        cam = st["cam"]
        kf = O.build_pyramid(orc, pcfg, cam, *st["frames"][0])
        O.make_keyframe(orc, kf)
        cur = O.build_pyramid(orc, pcfg, cam, *st["frames"][2])
        R0, T0 = np.eye(3), np.zeros(3)
        ref = orc.track_frames(kf, cur, R0, T0, cfg, N_LEVELS - 1, 0)
        base_evals.append(ref["evals"][:N_LEVELS])
        for l in range(N_LEVELS):
            ix, iy = project(cur.edges3d[l], kf.cams[l], ref["R"], ref["T"])
            w = kf.cams[l].w
            for name, kind, nbytes, bw, bh, _ in LAYOUTS:
                ln, sc = (lines_quad if kind == "quad" else lines_texel)(ix, iy, w, nbytes, bw, bh)
                fp[name][l] += (len(np.unique(ln)), len(np.unique(sc)), len(ix))
        exact_opt = kf.opt
        for name, mode in QUANT:
            kf.opt = [quantise(exact_opt[l], mode, cfg.edge_distance_lvl[l]) for l in range(N_LEVELS)]
            r = orc.track_frames(kf, cur, R0, T0, cfg, N_LEVELS - 1, 0)
            qres[name].append((rot_angle(r["R"] @ ref["R"].T), float(np.abs(r["T"] - ref["T"]).max()),
                               r["evals"][:N_LEVELS] == ref["evals"][:N_LEVELS], sum(r["evals"][:N_LEVELS])))
        kf.opt = exact_opt
        print("pair", s, ref["evals"][:N_LEVELS], flush=True)

    ev = np.mean(base_evals, axis=0)
    P(f"# Lookup-structure layout study for k_track (scratch/layout_study.py, CPU oracle, {n_pairs} synthetic VGA pairs, keyframe gap 2,")
    P(f"# {N_LEVELS} levels).  Mean evaluations per level {np.round(ev, 1).tolist()}, mean in-bounds points per level "
      f"{np.round(fp[LAYOUTS[0][0]][:, 2] / n_pairs).astype(int).tolist()}.")
    P("#")
    P("# Part 1: distinct 128-byte L2 lines / 32-byte sectors one evaluation of a pair touches (points projected at the")
    P("# converged pose).  'KB lines L0' is the L2 capacity one pair in flight occupies while it works on level 0; 'eval-weighted'")
    P("# sums lines x evaluations over the levels (proportional to the L2 lookups that can miss).")
    P("#")
    P(f"# {'layout':48s} {'loads/point':>12s} {'lines/pt L0':>11s} {'KB lines L0':>11s} {'KB sectors L0':>13s} {'eval-weighted lines (rel.)':>26s}")
    base_w = None
    for name, kind, nbytes, bw, bh, loads in LAYOUTS:
        a = fp[name] / n_pairs
        wsum = float((a[:, 0] * ev).sum())
        base_w = base_w or wsum
        P(f"  {name:48s} {loads:>12s} {a[0, 0] / a[0, 2]:11.3f} {a[0, 0] * 128 / 1024:11.0f} {a[0, 1] * 32 / 1024:13.0f} {wsum / base_w:26.2f}")
    P("#")
    P("# Part 2: oracle (float32) tracking with the keyframe structure quantised as a layout would store it, against the")
    P("# unquantised run of the same pair: pose difference, pairs whose per-level evaluation counts are unchanged, total")
    P("# evaluations.  (Parity bar of the path: 1e-4 rad / 1e-4 m after the same number of iterations.)")
    P("#")
    P(f"# {'structure':44s} {'max rot [rad]':>14s} {'median rot':>11s} {'max |dt| [m]':>13s} {'same evals':>11s} {'evals (exact: ' + str(int(np.sum(base_evals))) + ')':>22s}")
    for name, _ in QUANT:
        r = qres[name]
        rot = np.array([x[0] for x in r]); dt = np.array([x[1] for x in r])
        P(f"  {name:44s} {rot.max():14.2e} {np.median(rot):11.2e} {dt.max():13.2e} {sum(x[2] for x in r):>7d}/{len(r):<3d} {sum(x[3] for x in r):22d}")
    txt = "\n".join(out) + "\n"
    print(txt)
    with open(os.path.join(ROOT, "profiles", "r1_lookup_layout_study.txt"), "w") as f:
        f.write(txt)


if __name__ == "__main__":
    main()
