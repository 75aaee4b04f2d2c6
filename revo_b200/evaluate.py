"""Trajectory error metrics for trajectories in the TUM format (:mod:`revo_b200.tum_io`): absolute trajectory error after a
rigid alignment and relative pose error, the two numbers the reference's paper reports on the TUM RGB-D sequences
(BASELINE.md).  Host-side numpy; not part of the hot path.
"""
from __future__ import annotations

import numpy as np


def associate(ts_a, ts_b, max_dt: float = 0.02):
    """Greedy one-to-one association of two timestamp lists by closest time (|dt| < max_dt). Returns index pairs."""
    ts_a, ts_b = np.asarray(ts_a, np.float64), np.asarray(ts_b, np.float64)
    cand = [(abs(a - b), i, j) for i, a in enumerate(ts_a) for j, b in enumerate(ts_b) if abs(a - b) < max_dt] \
        if len(ts_a) * len(ts_b) < 4_000_000 else None
    if cand is None:      # long sequences: nearest neighbour through a sorted search
        order = np.argsort(ts_b)
        pos = np.clip(np.searchsorted(ts_b[order], ts_a), 1, len(ts_b) - 1)
        cand = []
        for i, p in enumerate(pos):
            for q in (p - 1, p):
                j = int(order[q])
                if abs(ts_a[i] - ts_b[j]) < max_dt:
                    cand.append((abs(ts_a[i] - ts_b[j]), i, j))
    cand.sort()
    used_a, used_b, out = set(), set(), []
    for _, i, j in cand:
        if i not in used_a and j not in used_b:
            used_a.add(i)
            used_b.add(j)
            out.append((i, j))
    out.sort()
    return out


def align_rigid(model_xyz, data_xyz):
    """Least-squares rigid transform (R, t) with R @ model + t ~= data (Horn / Kabsch via SVD). Inputs (n, 3)."""
    m, d = np.asarray(model_xyz, np.float64), np.asarray(data_xyz, np.float64)
    mc, dc = m.mean(axis=0), d.mean(axis=0)
    W = (d - dc).T @ (m - mc)
    U, _, Vt = np.linalg.svd(W)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    return R, dc - R @ mc


def ate(gt_poses, est_poses):
    """Absolute trajectory error of associated (n,4,4) pose lists after rigid alignment.
    Returns dict(rmse, mean, median, max, R, t)."""
    g = np.asarray(gt_poses, np.float64)[:, :3, 3]
    e = np.asarray(est_poses, np.float64)[:, :3, 3]
    R, t = align_rigid(e, g)
    err = np.linalg.norm((e @ R.T + t) - g, axis=1)
    return dict(rmse=float(np.sqrt((err ** 2).mean())), mean=float(err.mean()), median=float(np.median(err)),
                max=float(err.max()), R=R, t=t)


def rpe(gt_poses, est_poses, delta: int = 1):
    """Relative pose error over `delta` frames: E_i = (Q_i^-1 Q_{i+delta})^-1 (P_i^-1 P_{i+delta}).
    Returns dict(trans_rmse [m], rot_rmse [rad], n)."""
    Q, P = np.asarray(gt_poses, np.float64), np.asarray(est_poses, np.float64)
    tr, ro = [], []
    for i in range(len(Q) - delta):
        dq = np.linalg.inv(Q[i]) @ Q[i + delta]
        dp = np.linalg.inv(P[i]) @ P[i + delta]
        E = np.linalg.inv(dq) @ dp
        tr.append(np.linalg.norm(E[:3, 3]))
        ro.append(np.arccos(np.clip((np.trace(E[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)))
    tr, ro = np.asarray(tr), np.asarray(ro)
    return dict(trans_rmse=float(np.sqrt((tr ** 2).mean())), rot_rmse=float(np.sqrt((ro ** 2).mean())), n=len(tr))


def evaluate_files(gt_file: str, est_file: str, max_dt: float = 0.02):
    """ATE / RPE of an estimated TUM trajectory file against a ground-truth one."""
    from . import tum_io

    tg, pg = tum_io.read_trajectory(gt_file)
    te, pe = tum_io.read_trajectory(est_file)
    pairs = associate(tg, te, max_dt)
    if len(pairs) < 3:
        raise ValueError("fewer than 3 associated poses")
    ig, ie = [p[0] for p in pairs], [p[1] for p in pairs]
    return dict(n=len(pairs), ate=ate(pg[ig], pe[ie]), rpe=rpe(pg[ig], pe[ie]))
