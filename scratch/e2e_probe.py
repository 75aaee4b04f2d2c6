#!/usr/bin/env python
"""Where does the end-to-end step go?  Pinned host frames -> create_batch (upload on the copy stream + build) in isolation
and back to back, with the upload time from CUDA events on the copy stream."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from revo_b200 import api, synth

B, w, h = 128, 640, 480
bgr = torch.randint(0, 255, (4, B, h, w, 3), dtype=torch.uint8).pin_memory()
depth = (torch.rand((4, B, h, w), dtype=torch.float32) * 4 + 0.5).pin_memory()
fx, fy, cx, cy, _, _ = synth.intrinsics(w, h)
st = api.ImgPyramidSettings(PYR_MIN_LVL=3, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)
ctx = api.Context(0)
ctx.reserve(12 << 30)
# (1) one batch at a time
for i in range(4):
    t0 = time.perf_counter()
    b = api.PyramidBatch(ctx, st, bgr[i % 4], depth[i % 4], B)
    t1 = time.perf_counter()
    ctx.synchronize()
    t2 = time.perf_counter()
    print(f"single: enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms, upload {ctx.last_upload_ms():.2f} ms, build {ctx.last_timings()[0]:.2f} ms")
    b.destroy()
# (2) back to back, 8 batches, two in flight
ctx.synchronize()
t0 = time.perf_counter()
pend = []
for i in range(8):
    pend.append(api.PyramidBatch(ctx, st, bgr[i % 4], depth[i % 4], B))
    if len(pend) > 2:
        pend.pop(0).destroy()
    print(f"  enqueue {i}: t = {1e3*(time.perf_counter()-t0):.2f} ms")
ctx.synchronize()
t1 = time.perf_counter()
print(f"back to back: {1e3*(t1-t0)/8:.2f} ms per batch, last upload {ctx.last_upload_ms():.2f} ms")
# (3) raw copies with torch for reference
d_b = torch.empty_like(bgr[0], device="cuda"); d_d = torch.empty_like(depth[0], device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(8):
    d_b.copy_(bgr[i % 4], non_blocking=True); d_d.copy_(depth[i % 4], non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"torch copies: {1e3*(t1-t0)/8:.2f} ms per batch ({(bgr[0].numel()+depth[0].numel()*4)/((t1-t0)/8)/1e9:.1f} GB/s)")
