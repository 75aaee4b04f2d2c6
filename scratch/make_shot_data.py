#!/usr/bin/env python
"""Pre-renders a few synthetic VGA stream pairs (keyframe = frame 0, tracked frame = frame `gap`) on the CPU into
scratch/shot_data.npz, so that scratch/final_shot.py can run on a GPU box without torch (numpy + ctypes only).
8-bit BGR + raw 16-bit depth (1/5000 m), the dataset wire format."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from revo_b200 import synth  # noqa: E402

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gap = 2
w, h = 640, 480
bgr = np.empty((2, n_streams, h, w, 3), np.uint8)
d16 = np.empty((2, n_streams, h, w), np.uint16)
T = np.empty((2, n_streams, 4, 4), np.float64)
cam = None
for s in range(n_streams):
    st = synth.make_stream(2000 + s, gap + 1, w, h)
    cam = st["cam"]
    for j, f in enumerate((0, gap)):
        b, d = st["frames"][f]
        bgr[j, s] = b
        raw = np.round(d.astype(np.float64) * 5000.0)
        assert raw.max() < 65536
        d16[j, s] = raw.astype(np.uint16)
        assert np.array_equal(d16[j, s].astype(np.float32) * (np.float32(1.0) / np.float32(5000.0)), d)
        T[j, s] = st["T_w_c"][f]
    print("stream", s, flush=True)
np.savez_compressed(os.path.join(ROOT, "scratch", "shot_data.npz"), bgr=bgr, d16=d16, T=T, cam=np.asarray(cam, np.float64))
