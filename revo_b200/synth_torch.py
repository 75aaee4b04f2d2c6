"""Torch implementation of the synthetic scene renderer of ``synth.py`` (same scene model, same hash
texture, float64 geometry), so that bench.py can render thousands of distinct VGA frames in seconds on the
GPU (or, slowly, on the CPU).  Data generation only -- never inside a timed region.  The pixel noise and
depth holes come from a torch generator, so frames differ from the numpy renderer by the noise
realisation only (tests/test_gpu_bench_inputs.py checks the noise-free images agree)."""
from __future__ import annotations

import numpy as np
import torch

from . import synth

_M1 = -0x40A7B892E31B1A47  # 0xBF58476D1CE4E5B9 as int64
_M2 = -0x6B2FB644ECCEEE15  # 0x94D049BB133111EB as int64
_H0 = -0x61C8864680B583EB  # 0x9E3779B97F4A7C15 as int64


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right on int64"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def _hash_u32(*ints) -> torch.Tensor:
    h = None
    for a in ints:
        a = a.to(torch.int64)
        h = (a ^ _H0) if h is None else (h ^ a)
        h = h * _M1
        h = (h ^ _lsr(h, 31)) * _M2
        h = h ^ _lsr(h, 29)
    return _lsr(h, 32)


class TorchScene:
    def __init__(self, scene: synth.Scene, device):
        self.seed = scene.seed
        self.device = torch.device(device)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.boxes = torch.tensor(scene.boxes, **f64)
        self.room = torch.tensor(scene.room, **f64)
        self.cell = torch.tensor(scene.cell, **f64)
        self.tint = torch.tensor(scene.tint, **f64)

    @torch.no_grad()
    def render(self, T_wc: np.ndarray, cam, gen: torch.Generator | None, noise_sigma: float = 2.0, hole_frac: float = 0.02):
        """-> (bgr uint8 HxWx3, depth float32 HxW) tensors on self.device."""
        fx, fy, cx, cy, w, h = cam
        w, h = int(w), int(h)
        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)
        v, u = torch.meshgrid(torch.arange(h, **f64), torch.arange(w, **f64), indexing="ij")
        dc = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], dim=-1).reshape(-1, 3)
        R = torch.tensor(T_wc[:3, :3], **f64)
        o = torch.tensor(T_wc[:3, 3], **f64)
        d = dc @ R.T
        n = d.shape[0]
        inv = 1.0 / d
        t0 = (self.room[None, :3] - o[None, :]) * inv
        t1 = (self.room[None, 3:] - o[None, :]) * inv
        tfar = torch.maximum(t0, t1)
        best_t, ax = tfar.min(dim=1)
        sign = (d.gather(1, ax[:, None])[:, 0] > 0).to(torch.int64)
        best_face = ax * 2 + sign
        for b in range(self.boxes.shape[0]):
            box = self.boxes[b]
            t0 = (box[None, :3] - o[None, :]) * inv
            t1 = (box[None, 3:] - o[None, :]) * inv
            tn = torch.minimum(t0, t1)
            tf = torch.maximum(t0, t1)
            t_in, axn = tn.max(dim=1)
            t_out = tf.min(dim=1).values
            hit = (t_in < t_out) & (t_in > 0.05) & (t_in < best_t)
            sign = (d.gather(1, axn[:, None])[:, 0] > 0).to(torch.int64)
            best_t = torch.where(hit, t_in, best_t)
            best_face = torch.where(hit, 6 * (b + 1) + axn * 2 + sign, best_face)
        P = o[None, :] + d * best_t[:, None]
        axis = (best_face % 6) // 2
        a0 = torch.where(axis == 0, P[:, 1], P[:, 0])
        a1 = torch.where(axis == 2, P[:, 1], P[:, 2])
        cs = self.cell[best_face]
        iu = torch.floor(a0 / cs[:, 0])
        iv = torch.floor(a1 / cs[:, 1])
        seed = torch.full_like(best_face, self.seed)
        hsh = _hash_u32(best_face, iu, iv, seed)
        albedo = 30.0 + (hsh % 201).to(torch.float64)
        hs2 = _hash_u32(best_face + 1000, torch.floor(iu / 3), torch.floor(iv / 2), seed)
        alb2 = 30.0 + (_lsr(hs2, 8) % 201).to(torch.float64)
        albedo = torch.where((hs2 % 100) < 35, alb2, albedo)
        img = albedo[:, None] * self.tint[best_face]
        if gen is not None and noise_sigma > 0:
            img = img + torch.randn((n, 3), generator=gen, **f64) * noise_sigma
        bgr = torch.clamp(torch.round(img), 0, 255).to(torch.uint8).reshape(h, w, 3)
        z16 = torch.clamp(torch.round(best_t * synth.DEPTH_SCALE), 0, 65535)
        if gen is not None and hole_frac > 0:
            z16 = torch.where(torch.rand((n,), generator=gen, **f64) < hole_frac, torch.zeros_like(z16), z16)
        scale = float(np.float64(np.float32(1.0) / np.float32(synth.DEPTH_SCALE)))
        depth = (z16 * scale).to(torch.float32).reshape(h, w)
        return bgr, depth


def make_stream_poses(seed: int, n_frames: int, max_trans=0.015, max_rot_deg=0.8):
    """The camera trajectory of synth.make_stream (same rng draws)."""
    rng = np.random.default_rng([seed, 17])
    T = synth.base_pose(seed)
    vel = np.zeros(6)
    poses = []
    for _ in range(n_frames):
        poses.append(T.copy())
        vel = 0.85 * vel + 0.15 * np.concatenate([rng.normal(0, max_trans, 3), rng.normal(0, np.deg2rad(max_rot_deg), 3)])
        nt, nr = np.linalg.norm(vel[:3]), np.linalg.norm(vel[3:])
        if nt > max_trans:
            vel[:3] *= max_trans / nt
        if nr > np.deg2rad(max_rot_deg):
            vel[3:] *= np.deg2rad(max_rot_deg) / nr
        T = T @ synth.se3_exp(vel)
    return poses


@torch.no_grad()
def render_streams(seeds, n_frames: int, w: int, h: int, device, out_bgr: torch.Tensor, out_depth: torch.Tensor):
    """Render len(seeds) independent streams of n_frames each into
    out_bgr[frame, stream] (n_frames, S, h, w, 3) u8 and out_depth[frame, stream] (n_frames, S, h, w) f32
    (any device; typically pinned host memory).  Returns the ground-truth poses [stream][frame] (4x4)."""
    cam = synth.intrinsics(w, h)
    all_poses = []
    for si, seed in enumerate(seeds):
        sc = TorchScene(synth.Scene.make(int(seed)), device)
        poses = make_stream_poses(int(seed), n_frames)
        gen = torch.Generator(device=device)
        gen.manual_seed(int(seed) * 7919 + 13)
        for fi in range(n_frames):
            bgr, depth = sc.render(poses[fi], cam, gen)
            out_bgr[fi, si].copy_(bgr, non_blocking=True)
            out_depth[fi, si].copy_(depth, non_blocking=True)
        all_poses.append(poses)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    return cam, all_poses
