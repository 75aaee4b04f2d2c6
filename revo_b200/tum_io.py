"""Dataset wire formats either side of the hot path: the TUM RGB-D association list the reference's reader
walks (io/iowrapperRGBD.cpp:301-333) and the trajectory file REVO::writePose produces (system/system.cpp:75-79).
Host-side only; images are handed to the device as they are on disk -- 8-bit BGR and RAW 16-bit depth
(``revo_pyr_create_batch_u16`` converts with 1/DEPTH_SCALE_FACTOR on the device, the reader's ``convertTo``).
"""
from __future__ import annotations

import os
from typing import Iterator, List, Tuple

import numpy as np


def read_associations(path: str, skip_first_n: int = 0) -> List[Tuple[float, str, float, str]]:
    """Lines ``rgb_ts rgb_file depth_ts depth_file``; '#' comments and empty lines ignored; the first ``skip_first_n`` data
    lines are skipped (SKIP_FIRST_N_FRAMES, iowrapperRGBD.cpp:310-316)."""
    out = []
    n = 0
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            n += 1
            if n <= skip_first_n:
                continue
            tok = line.split()
            if len(tok) < 4:
                raise ValueError(f"{path}: malformed association line: {line!r}")
            out.append((float(tok[0]), tok[1], float(tok[2]), tok[3]))
    return out


def load_frame(folder: str, rgb_file: str, depth_file: str):
    """(bgr uint8 HxWx3, raw depth uint16 HxW) exactly as cv::imread(rgb) / cv::imread(depth, UNCHANGED) return them."""
    import cv2

    bgr = cv2.imread(os.path.join(folder, rgb_file), cv2.IMREAD_COLOR)
    raw = cv2.imread(os.path.join(folder, depth_file), cv2.IMREAD_UNCHANGED)
    if bgr is None or raw is None:
        raise FileNotFoundError(f"cannot read {rgb_file} / {depth_file} under {folder}")
    if raw.dtype != np.uint16 or raw.ndim != 2:
        raise ValueError(f"{depth_file}: expected a 16-bit single-channel depth image, got {raw.dtype} {raw.shape}")
    return bgr, raw


def iter_frames(folder: str, assoc_file: str = "associate.txt", skip_first_n: int = 0,
                use_depth_timestamp: bool = False) -> Iterator[Tuple[float, np.ndarray, np.ndarray]]:
    """Yields (timestamp, bgr, raw depth) in file order.  The timestamp a pyramid is stamped with is
    ``useDepthTimeStamp ? depthTimeStamp : rgbTimeStamp`` (iowrapperRGBD.cpp:266); both dataset configurations the reference
    ships (dataset_tum1.yaml, orbbec_dataset.yaml) set ``useDepthTimeStamp: 0``, so the rgb timestamp is the default here."""
    for rgb_ts, rgb_file, depth_ts, depth_file in read_associations(os.path.join(folder, assoc_file), skip_first_n):
        bgr, raw = load_frame(folder, rgb_file, depth_file)
        yield (depth_ts if use_depth_timestamp else rgb_ts), bgr, raw


def quaternion_from_R(R) -> np.ndarray:
    """Eigen::Quaternionf(R) (x, y, z, w): Shepperd's method with Eigen's branch order."""
    R = np.asarray(R, np.float64).reshape(3, 3)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def pose_to_tum_string(T_w_c, timestamp: float) -> str:
    """``timestamp tx ty tz qx qy qz qw`` as REVO::writePose formats it (std::fixed: 6 decimals for the timestamp, then
    setprecision(9))."""
    T = np.asarray(T_w_c, np.float64).reshape(4, 4)
    q = quaternion_from_R(T[:3, :3])
    return f"{timestamp:.6f} " + " ".join(f"{v:.9f}" for v in (T[0, 3], T[1, 3], T[2, 3], q[0], q[1], q[2], q[3]))


def write_trajectory(path: str, timestamps, poses_w_c) -> None:
    with open(path, "w") as f:
        for ts, T in zip(timestamps, poses_w_c):
            f.write(pose_to_tum_string(T, ts) + "\n")


def read_trajectory(path: str):
    """-> (timestamps (n,), poses (n,4,4)) from a TUM trajectory file."""
    ts, poses = [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            v = [float(x) for x in line.split()]
            x, y, z, w = v[4:8]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            T = np.eye(4)
            T[:3, :3] = R
            T[:3, 3] = v[1:4]
            ts.append(v[0])
            poses.append(T)
    return np.array(ts), np.array(poses)
