"""TUM wire formats (revo_b200/tum_io.py): association list parsing like the reference's reader
(io/iowrapperRGBD.cpp:301-333), 8-bit colour + raw 16-bit depth round trip through PNG files, and the trajectory format of
REVO::writePose (system/system.cpp:75-79)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_association_list_and_frames(tmp_path):
    import cv2

    from revo_b200 import synth, tum_io

    p = synth.make_pair(4, 160, 120)
    os.makedirs(tmp_path / "rgb")
    os.makedirs(tmp_path / "depth")
    lines = ["# comment", ""]
    raws = []
    for i, (bgr, depth) in enumerate((p["key"], p["cur"])):
        raw = np.round(depth.astype(np.float64) * 5000.0).astype(np.uint16)
        raws.append(raw)
        cv2.imwrite(str(tmp_path / "rgb" / f"{i}.png"), bgr)
        cv2.imwrite(str(tmp_path / "depth" / f"{i}.png"), raw)
        lines.append(f"{1305031102.175304 + i:.6f} rgb/{i}.png {1305031102.160407 + i:.6f} depth/{i}.png")
    (tmp_path / "associate.txt").write_text("\n".join(lines) + "\n")
    assoc = tum_io.read_associations(str(tmp_path / "associate.txt"))
    assert len(assoc) == 2 and assoc[0][1] == "rgb/0.png" and assoc[1][3] == "depth/1.png"
    assert len(tum_io.read_associations(str(tmp_path / "associate.txt"), skip_first_n=1)) == 1
    frames = list(tum_io.iter_frames(str(tmp_path)))
    # useDepthTimeStamp: 0 in the reference's dataset configurations -> the rgb timestamp (iowrapperRGBD.cpp:266)
    assert len(frames) == 2 and abs(frames[1][0] - 1305031103.175304) < 1e-6
    assert abs(next(tum_io.iter_frames(str(tmp_path), use_depth_timestamp=True))[0] - 1305031102.160407) < 1e-6
    for (ts, bgr, raw), ref_raw, (ref_bgr, ref_depth) in zip(frames, raws, (p["key"], p["cur"])):
        assert bgr.dtype == np.uint8 and np.array_equal(bgr, ref_bgr)
        assert raw.dtype == np.uint16 and np.array_equal(raw, ref_raw)
        # the reader's conversion reproduces the float depth the synthetic scene was quantised from
        assert np.array_equal(raw.astype(np.float32) * (np.float32(1.0) / np.float32(5000.0)), ref_depth)


def test_trajectory_round_trip(tmp_path):
    from revo_b200 import synth, tum_io

    rng = np.random.default_rng(1)
    poses = [synth.se3_exp(rng.normal(0, 0.5, 6)) for _ in range(5)]
    poses.append(np.diag([-1.0, -1.0, 1.0, 1.0]))            # trace <= 0: the other branch of the quaternion conversion
    ts = [1305031102.175304 + 0.033 * i for i in range(len(poses))]
    tum_io.write_trajectory(str(tmp_path / "traj.txt"), ts, poses)
    first = (tmp_path / "traj.txt").read_text().splitlines()[0].split()
    assert len(first) == 8 and first[0] == "1305031102.175304" and all(len(v.split(".")[1]) == 9 for v in first[1:])
    t2, p2 = tum_io.read_trajectory(str(tmp_path / "traj.txt"))
    assert np.allclose(t2, ts, atol=1e-6)
    for a, b in zip(poses, p2):
        assert np.allclose(a, b, atol=2e-8 * 10)
    q = tum_io.quaternion_from_R(np.eye(3))
    assert np.array_equal(q, [0, 0, 0, 1])


def test_ate_rpe_on_transformed_noisy_trajectory(tmp_path):
    """ATE is invariant to a rigid transform of the estimate and measures the added noise; RPE of an exact copy is 0."""
    from revo_b200 import evaluate, synth, tum_io

    rng = np.random.default_rng(3)
    gt = [np.eye(4)]
    for _ in range(59):
        gt.append(gt[-1] @ synth.se3_exp(np.r_[rng.normal(0, 0.01, 3), rng.normal(0, 0.01, 3)]))
    gt = np.array(gt)
    ts = 100.0 + 0.033 * np.arange(len(gt))
    G = synth.se3_exp(np.array([0.3, -0.2, 0.5, 0.2, -0.4, 0.9]))
    est = np.array([G @ T for T in gt])
    a = evaluate.ate(gt, est)
    assert a["rmse"] < 1e-9
    assert evaluate.rpe(gt, est)["trans_rmse"] < 1e-9 and evaluate.rpe(gt, est, delta=5)["rot_rmse"] < 1e-6
    noisy = est.copy()
    noisy[:, :3, 3] += rng.normal(0, 0.01, (len(gt), 3))
    a = evaluate.ate(gt, noisy)
    assert 0.01 < a["rmse"] < 0.025                       # sigma * sqrt(3) = 0.017
    # through the files, with the estimate's timestamps shifted by a few milliseconds and one frame dropped
    tum_io.write_trajectory(str(tmp_path / "gt.txt"), ts, gt)
    keep = [i for i in range(len(gt)) if i != 7]
    tum_io.write_trajectory(str(tmp_path / "est.txt"), ts[keep] + 0.004, noisy[keep])
    r = evaluate.evaluate_files(str(tmp_path / "gt.txt"), str(tmp_path / "est.txt"))
    assert r["n"] == len(gt) - 1 and abs(r["ate"]["rmse"] - a["rmse"]) < 2e-3
