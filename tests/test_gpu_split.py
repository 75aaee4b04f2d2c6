"""Multi-GPU edge split of one pair (needs >= 2 GPUs: run under `gpurun --gpus 2`).  All ranks must return the
bit-identical pose (they sum the same partials in the same rank order), equal to the single-GPU pose up to
summation order, and a second launch must work on the same mailboxes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rot_angle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("w,h", [(640, 480)])
def test_edge_split_two_gpus(tmp_path, w, h):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29551")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29551", os.path.join(ROOT, "tests", "_split_worker.py"), str(tmp_path), str(w), str(h)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    a, b = [np.load(tmp_path / f"split{k}.npz") for k in range(2)]
    assert np.array_equal(a["R2"], b["R2"]) and np.array_equal(a["T2"], b["T2"])          # ranks agree bit for bit
    assert np.array_equal(a["R2"], a["R3"]) and np.array_equal(a["T2"], a["T3"])          # second launch identical
    assert rot_angle(a["R1"], a["R2"]) <= 2e-4 and np.linalg.norm(a["T1"] - a["T2"]) <= 2e-4
    assert list(a["s"]) == list(b["s"])
    print("evals single", a["ev1"], "split", a["ev2"])
