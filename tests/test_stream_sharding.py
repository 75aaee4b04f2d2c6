"""Host logic of the multi-stream driver and of the N>1 sharding (CPU, gloo, world_size 2):
 * StreamTracker (motion-model init, keyframe promotion, world-pose composition: system/system.cpp:191-271)
   follows the renderer's ground-truth trajectory with the oracle backend;
 * two gloo ranks that each own half of the streams produce exactly the poses of the single-process run
   (streams are independent: no data-path collective), and the timing reduction is a MAX over ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _single_process(seeds, n_frames, policy="interval"):
    sys.path.insert(0, ROOT)
    from oracle.stream_backend import OracleBackend
    from revo_b200 import synth
    from revo_b200.stream import StreamTracker

    w, h = 160, 120
    cam = synth.intrinsics(w, h)
    fast = policy == "vote"
    streams = [synth.make_stream(sd, n_frames, w, h, max_trans=0.01 + 0.04 * (fast and sd % 2), max_rot_deg=0.5 + 2.5 * (fast and sd % 2))
               for sd in seeds]
    st = StreamTracker(OracleBackend(cam, 3), len(seeds), kf_interval=3, kf_policy=policy)
    frame = lambda i: (np.stack([s["frames"][i][0] for s in streams]), np.stack([s["frames"][i][1] for s in streams]))
    st.start(*frame(0))
    for i in range(1, n_frames):
        st.step(*frame(i))
    return st, streams


def test_stream_tracker_follows_ground_truth():
    from revo_b200 import synth

    st, streams = _single_process([300, 301], 6)
    for s in range(2):
        T_gt = np.linalg.inv(streams[s]["T_w_c"][0]) @ streams[s]["T_w_c"][5]
        D = np.linalg.inv(T_gt) @ st.T_w_c[s].astype(np.float64)
        assert synth.rot_angle(D[:3, :3]) < 6e-3 and np.linalg.norm(D[:3, 3]) < 1.5e-2
    assert st.total_evals > 0 and st.frame == 5


def test_two_rank_gloo_sharding_equals_single_process(tmp_path):
    B, n_frames = 2, 5
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "_shard_worker.py"), str(tmp_path), str(B), str(n_frames)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ranks = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert ranks[0]["dt"] == ranks[1]["dt"]                       # MAX over ranks: identical on every rank
    assert ranks[0]["frames"][0] == 2 * B * (n_frames - 1)        # whole-job units
    st, _ = _single_process([300, 301, 302, 303], n_frames)
    both = np.concatenate([ranks[0]["T_w_c"], ranks[1]["T_w_c"]])
    assert np.array_equal(both, st.T_w_c)                        # same streams, same poses, bit for bit
    assert int(ranks[0]["evals"]) + int(ranks[1]["evals"]) == st.total_evals


def test_two_rank_gloo_sharding_with_the_vote_policy(tmp_path):
    """The same sharding with the reference's keyframe policy (StreamTracker(kf_policy="vote")): every rank votes, promotes and
    re-aligns its own streams only -- the keyframe switches of the fast streams happen on whichever rank owns them, and the
    sharded job reproduces the single-process poses bit for bit."""
    B, n_frames = 2, 8
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29543", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "_shard_worker.py"), str(tmp_path), str(B), str(n_frames), "vote"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ranks = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    st, _ = _single_process([300, 301, 302, 303], n_frames, policy="vote")
    assert np.array_equal(np.concatenate([ranks[0]["T_w_c"], ranks[1]["T_w_c"]]), st.T_w_c)
    assert int(ranks[0]["retracks"]) + int(ranks[1]["retracks"]) == st.n_retracks >= 1
    assert int(ranks[0]["evals"]) + int(ranks[1]["evals"]) == st.total_evals


def test_pipelined_stepping_equals_sequential():
    """step_pipelined() with two frames in flight (prefetch, prefetch, then one new frame per step) tracks the same frames
    in the same order as step(): identical poses with the (deterministic) oracle backend, and close() releases what is
    still in flight."""
    sys.path.insert(0, ROOT)
    from oracle.stream_backend import OracleBackend
    from revo_b200 import synth
    from revo_b200.stream import StreamTracker

    w, h, n_frames = 160, 120, 6
    cam = synth.intrinsics(w, h)
    streams = [synth.make_stream(sd, n_frames, w, h) for sd in (310, 311)]
    frame = lambda i: (np.stack([s["frames"][i][0] for s in streams]), np.stack([s["frames"][i][1] for s in streams]))

    class CountingBackend(OracleBackend):
        created = destroyed = 0

        def create(self, bgr, depth, n):
            CountingBackend.created += 1
            return super().create(bgr, depth, n)

        def destroy(self, handles):
            CountingBackend.destroyed += 1

    seq = StreamTracker(OracleBackend(cam, 3), 2, kf_interval=3)
    seq.start(*frame(0))
    for i in range(1, n_frames):
        seq.step(*frame(i))

    pip = StreamTracker(CountingBackend(cam, 3), 2, kf_interval=3)
    pip.start(*frame(0))
    pip.prefetch(*frame(1))
    pip.prefetch(*frame(2))
    for i in range(1, n_frames):
        nxt = frame(i + 2) if i + 2 < n_frames else (None, None)
        pip.step_pipelined(*nxt)
    assert pip.frame == seq.frame == n_frames - 1
    assert np.array_equal(pip.T_w_c, seq.T_w_c) and pip.total_evals == seq.total_evals
    assert len(pip._pending) == 0
    pip.close()
    assert CountingBackend.created == n_frames and CountingBackend.destroyed >= 3
