"""Seeded synthetic RGB-D scenes: an analytically ray-cast box
room with floating boxes, piecewise-constant random-rectangle albedo, sigma=2
pixel noise, TUM-like depth (1/5000 m quantisation, 2 % holes), rendered from an
exact SE(3) camera pose so the ground-truth relative pose of any two frames is
known.  Pure numpy; shared by tests, bench.py and the CPU-baseline leg so both
arms see byte-identical inputs.

Input format mirrors what ``IOWrapperRGBD::readNextFrame`` hands to the pyramid
constructor (io/iowrapperRGBD.cpp:262-327): ``bgr`` uint8 HxWx3 (OpenCV BGR
order) and ``depth`` float32 HxW in metres, 0 = invalid.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# config/dataset_tum1.yaml:8-11 (TUM fr1 intrinsics, VGA)
TUM_FR1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, w=640, h=480)
DEPTH_SCALE = 5000.0  # config/dataset_tum1.yaml:45


def intrinsics(w: int, h: int):
    """(fx,fy,cx,cy,w,h): TUM fr1 scaled to the width for 4:3 sizes, else the
    1080p set of SURVEY 8(d)."""
    if w * 3 == h * 4:
        s = w / 640.0
        return (TUM_FR1["fx"] * s, TUM_FR1["fy"] * s, TUM_FR1["cx"] * s, TUM_FR1["cy"] * s, w, h)
    f = 1050.0 * w / 1920.0
    return (f, f, (w - 1) / 2.0, (h - 1) / 2.0, w, h)


# ---------------------------------------------------------------------------
# SE(3) helpers (float64, closed form; independent of the oracle and the product)
# ---------------------------------------------------------------------------
def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def se3_exp(xi):
    """xi = (upsilon, omega) -> 4x4, same convention as Sophus::SE3::exp."""
    xi = np.asarray(xi, np.float64)
    u, w = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    W = hat(w)
    if th < 1e-10:
        R = np.eye(3) + W
        V = np.eye(3) + 0.5 * W
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th**2 * (W @ W)
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * (W @ W)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ u
    return T


def rot_angle(R):
    """Rotation angle from the skew part (well conditioned for small angles, unlike arccos(trace))."""
    R = np.asarray(R, np.float64)
    s = 0.5 * np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arctan2(s, 0.5 * (np.trace(R) - 1.0)))


# ---------------------------------------------------------------------------
# Scene
# ---------------------------------------------------------------------------
def _hash_u32(*ints):
    """Vectorised integer hash -> uint32 (splitmix-style)."""
    h = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        for a in ints:
            a = np.asarray(a).astype(np.int64).astype(np.uint64)
            h = (h ^ a) * np.uint64(0xBF58476D1CE4E5B9)
            h = (h ^ (h >> np.uint64(31))) * np.uint64(0x94D049BB133111EB)
            h = h ^ (h >> np.uint64(29))
    return (h >> np.uint64(32)).astype(np.uint32)


@dataclass
class Scene:
    seed: int
    boxes: np.ndarray      # (nb, 6) xmin,ymin,zmin,xmax,ymax,zmax  (world)
    room: np.ndarray       # (6,) xmin,ymin,zmin,xmax,ymax,zmax ; camera looks along +z, zmin wall is behind (unused)
    cell: np.ndarray       # (n_faces, 2) texture cell size in metres
    tint: np.ndarray       # (n_faces, 3) BGR tint

    @staticmethod
    def make(seed: int, n_boxes: int = 6, cell_range=(0.05, 0.30)) -> "Scene":
        rng = np.random.default_rng(seed)
        room = np.array([-2.6, -1.6, -1.0, 2.6, 1.6, 4.5])
        boxes = []
        for _ in range(n_boxes):
            c = np.array([rng.uniform(-1.6, 1.6), rng.uniform(-1.0, 1.0), rng.uniform(1.3, 3.4)])
            s = rng.uniform(0.18, 0.55, size=3)
            boxes.append(np.concatenate([c - s, c + s]))
        boxes = np.array(boxes)
        n_faces = 6 * (n_boxes + 1)
        cell = rng.uniform(cell_range[0], cell_range[1], size=(n_faces, 2))
        tint = rng.uniform(0.85, 1.0, size=(n_faces, 3))
        return Scene(seed, boxes, room, cell, tint)

    # -- rendering ---------------------------------------------------------
    def render(self, T_wc: np.ndarray, cam, noise_seed: int, noise_sigma: float = 2.0, hole_frac: float = 0.02):
        """Render from camera-to-world pose T_wc. Returns (bgr u8 HxWx3, depth f32 HxW metres)."""
        fx, fy, cx, cy, w, h = cam
        w, h = int(w), int(h)
        u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        dc = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1).reshape(-1, 3)
        R, o = T_wc[:3, :3], T_wc[:3, 3]
        d = dc @ R.T                       # world ray direction, depth == ray parameter (d_c.z == 1)
        n = d.shape[0]
        best_t = np.full(n, np.inf)
        best_face = np.zeros(n, np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            # room: inside-out hit = the exit point of the slab test
            t0 = (self.room[None, :3] - o[None, :]) * inv
            t1 = (self.room[None, 3:] - o[None, :]) * inv
            tfar = np.maximum(t0, t1)
            ax = np.argmin(tfar, axis=1)
            t_exit = tfar[np.arange(n), ax]
            sign = (d[np.arange(n), ax] > 0).astype(np.int64)
            best_t = t_exit
            best_face = ax * 2 + sign
            for b, box in enumerate(self.boxes):
                t0 = (box[None, :3] - o[None, :]) * inv
                t1 = (box[None, 3:] - o[None, :]) * inv
                tn = np.minimum(t0, t1)
                tf = np.maximum(t0, t1)
                axn = np.argmax(tn, axis=1)
                t_in = tn[np.arange(n), axn]
                t_out = tf.min(axis=1)
                hit = (t_in < t_out) & (t_in > 0.05) & (t_in < best_t)
                sign = (d[np.arange(n), axn] > 0).astype(np.int64)
                best_t = np.where(hit, t_in, best_t)
                best_face = np.where(hit, 6 * (b + 1) + axn * 2 + sign, best_face)
        P = o[None, :] + d * best_t[:, None]
        axis = (best_face % 6) // 2
        # the two in-plane coordinates of the hit face
        a0 = np.where(axis == 0, P[:, 1], P[:, 0])
        a1 = np.where(axis == 2, P[:, 1], P[:, 2])
        cs = self.cell[best_face]
        iu = np.floor(a0 / cs[:, 0])
        iv = np.floor(a1 / cs[:, 1])
        hsh = _hash_u32(best_face, iu, iv, self.seed)
        albedo = 30.0 + (hsh % np.uint32(201)).astype(np.float64)          # 30..230
        # second, coarser layer that overrides ~35 % of the cells -> rectangles of mixed size
        hs2 = _hash_u32(best_face + 1000, np.floor(iu / 3), np.floor(iv / 2), self.seed)
        albedo = np.where((hs2 % np.uint32(100)) < 35, 30.0 + ((hs2 >> np.uint32(8)) % np.uint32(201)), albedo)
        rng = np.random.default_rng([self.seed, noise_seed, 7])
        img = albedo[:, None] * self.tint[best_face] + rng.normal(0.0, noise_sigma, size=(n, 3))
        bgr = np.clip(np.rint(img), 0, 255).astype(np.uint8).reshape(h, w, 3)
        z16 = np.clip(np.rint(best_t * DEPTH_SCALE), 0, 65535).astype(np.uint16)
        holes = rng.random(n) < hole_frac
        z16[holes] = 0
        # depth.convertTo(CV_32FC1, 1.0f/DEPTH_SCALE_FACTOR)  (io/iowrapperRGBD.cpp:325-327)
        scale = np.float64(np.float32(1.0) / np.float32(DEPTH_SCALE))
        depth = (z16.astype(np.float64) * scale).astype(np.float32).reshape(h, w)
        return bgr, depth


def base_pose(seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, 11])
    xi = np.concatenate([rng.uniform(-0.15, 0.15, 3), rng.uniform(-0.06, 0.06, 3)])
    return se3_exp(xi)


# motion of BASELINE.md section 3, config 1
XI_CONFIG1 = np.array([0.010, -0.006, 0.008, 0.004, -0.006, 0.003])


def make_pair(seed: int, w: int = 640, h: int = 480, xi=None, max_trans=0.02, max_rot_deg=1.0):
    """One key/current frame pair.  Returns dict(cam, key=(bgr,depth), cur=(bgr,depth), T_kf_cur 4x4, xi).
    ``p_kf = R p_cur + T`` with (R,T) = T_kf_cur, the direction convention of
    system/system.cpp:191-192."""
    cam = intrinsics(w, h)
    scene = Scene.make(seed)
    T_w_kf = base_pose(seed)
    if xi is None:
        rng = np.random.default_rng([seed, 13])
        dirt = rng.normal(size=3)
        dirr = rng.normal(size=3)
        xi = np.concatenate([dirt / np.linalg.norm(dirt) * rng.uniform(0.3, 1.0) * max_trans,
                             dirr / np.linalg.norm(dirr) * rng.uniform(0.3, 1.0) * np.deg2rad(max_rot_deg)])
    T_kf_cur = se3_exp(xi)
    T_w_cur = T_w_kf @ T_kf_cur
    key = scene.render(T_w_kf, cam, noise_seed=0)
    cur = scene.render(T_w_cur, cam, noise_seed=1)
    return dict(cam=cam, key=key, cur=cur, T_kf_cur=T_kf_cur, xi=np.asarray(xi, np.float64))


def make_stream(seed: int, n_frames: int, w: int = 640, h: int = 480, max_trans=0.015, max_rot_deg=0.8):
    """TUM-fr1-style stream: smooth random-walk velocity, <=1.5 cm and <=0.8 deg per frame.
    Returns dict(cam, frames=[(bgr,depth)], T_w_c=[4x4])."""
    cam = intrinsics(w, h)
    scene = Scene.make(seed)
    rng = np.random.default_rng([seed, 17])
    T = base_pose(seed)
    vel = np.zeros(6)
    frames, poses = [], []
    for i in range(n_frames):
        frames.append(scene.render(T, cam, noise_seed=i))
        poses.append(T.copy())
        vel = 0.85 * vel + 0.15 * np.concatenate([rng.normal(0, max_trans, 3), rng.normal(0, np.deg2rad(max_rot_deg), 3)])
        nt, nr = np.linalg.norm(vel[:3]), np.linalg.norm(vel[3:])
        if nt > max_trans:
            vel[:3] *= max_trans / nt
        if nr > np.deg2rad(max_rot_deg):
            vel[3:] *= np.deg2rad(max_rot_deg) / nr
        T = T @ se3_exp(vel)
    return dict(cam=cam, frames=frames, T_w_c=poses)
