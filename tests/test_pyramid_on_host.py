"""The whole pyramid construction on the CPU: every kernel of pyramid.cu and of the bit-mask Canny, with their launch code and
the launch sequence of ``capi.cu:create_batch_impl`` / ``launch_keyframe``, compiled from the CUDA source text against the
emulation layer of ``tests/_cuda_emu.py``, against the oracle pyramid (cv2 + the C restatement): gray, depth pyramid, Canny,
patch histogram, edge fill-in, 3-D edge list (tile-major list as a set, reference-order list exactly), exact EDT and the
lookup structure -- ``test_pyramid_bit_exact`` of the `-m gpu` suite without a GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import synth_pair

# an emulation deadlock must not hang the suite (the C call cannot be interrupted by a signal: kill the run instead)
pytestmark = pytest.mark.timeout(1200, method="thread")


class EmuLevelOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("gray", "depth", "edges", "edges_orig", "hist", "pts", "n_pts", "nz_patches", "dt", "opt",
                                          "pts_ref", "n_ref", "opt_f4")] + \
               [(n, C.c_int) for n in ("w", "h", "patch", "cap", "n_tiles")] + [(n, C.c_float) for n in ("fx", "fy", "cx", "cy")]


@pytest.fixture(scope="module")
def pyr_emu(tmp_path_factory):
    import _cuda_emu

    return _cuda_emu.build_pyramid(str(tmp_path_factory.mktemp("pyramid_emu")))


def run_pyramid(lib, bgr, depth, cam, n_levels, n_frames=1, keyframe=True, n_percentage=0.3):
    fx, fy, cx, cy, w0, h0 = cam
    outs = (EmuLevelOut * n_levels)()
    arrays = []
    for l in range(n_levels):
        w, h = w0 >> l, h0 >> l
        patch = max(1, 20 >> l)
        cap = w * h // 2 + 1024
        a = dict(gray=np.zeros((h, w), np.uint8), depth=np.zeros((h, w), np.float32), edges=np.zeros((h, w), np.uint8),
                 edges_orig=np.zeros((h, w), np.uint8), hist=np.zeros((max(1, h // patch), max(1, w // patch)), np.uint8),
                 pts=np.zeros((cap, 4), np.float32), n_pts=np.zeros(2, np.int32), nz_patches=np.zeros(2, np.int32),
                 dt=np.zeros((h, w), np.float32), opt=np.zeros((((h + 3) // 4) * ((w + 3) // 4) * 16, 2), np.uint32), pts_ref=np.zeros((w * h, 4), np.float32),
                 n_ref=np.zeros(2, np.int32), opt_f4=np.zeros((h, w, 4), np.float32))
        arrays.append(a)
        o = outs[l]
        for k, v in a.items():
            setattr(o, k, v.ctypes.data)
        o.w, o.h, o.patch, o.cap, o.n_tiles = w, h, patch, cap, ((w + 7) // 8) * ((h + 3) // 4)
        s = 2.0 ** -l                                    # level_camera: camerapyr.h:98-103
        o.fx, o.fy, o.cx, o.cy = fx * s, fy * s, cx * s, cy * s
    bgr = np.ascontiguousarray(bgr, np.uint8)
    depth = np.ascontiguousarray(depth, np.float32)
    rc = lib.emu_pyramid(bgr.ctypes.data_as(C.c_void_p), C.c_int(bgr.shape[2]), depth.ctypes.data_as(C.c_void_p), C.c_int(n_frames),
                         C.c_int(n_levels), outs, C.c_int(100 * 100), C.c_int(150 * 150), C.c_float(0.1), C.c_float(5.2), C.c_int(1),
                         C.c_float(n_percentage), C.c_int(int(keyframe)))
    assert rc == 0
    return arrays


def compare(arrays, po, n_levels, keyframe=True):
    for l in range(n_levels):
        a = arrays[l]
        assert np.array_equal(a["gray"], po.gray[l]), l
        assert np.array_equal(a["depth"].view(np.uint32), np.asarray(po.depth[l], np.float32).view(np.uint32)), l
        assert np.array_equal(a["edges_orig"], po.edges_orig[l]), l
        assert np.array_equal(a["edges"], po.edges[l]), l
        if po.hist[l] is not None and l < 3:
            hh, hw = np.asarray(po.hist[l]).shape
            assert np.array_equal(a["hist"][:hh, :hw], po.hist[l]), l
        want = np.asarray(po.edges3d[l], np.float32).reshape(-1, 4)
        n = int(a["n_ref"][0])
        assert n == len(want) == int(a["n_pts"][0]), (l, n, len(want), a["n_pts"][0])
        assert np.array_equal(a["pts_ref"][:n].view(np.uint32), want.view(np.uint32)), l          # reference (column-major) order
        got_set = {tuple(r) for r in a["pts"][:n].view(np.uint32).tolist()}                       # device (tile-major) order: same set
        assert got_set == {tuple(r) for r in want.view(np.uint32).tolist()}, l
        if keyframe:
            assert np.array_equal(a["dt"].view(np.uint32), np.asarray(po.dt[l], np.float32).view(np.uint32)), l
            ow = np.asarray(po.opt[l], np.float32)
            assert np.array_equal(a["opt_f4"][1:-1, :, :3].view(np.uint32), ow[1:-1, :, :3].view(np.uint32)), l
            # device layout: 8-byte texels {dt float32 bits | snorm16 gx, gy with step 1/32764} in 4x4 tiles (internal.h: opt_texel_index);
            # dt is 0 in rows 0 and h-1 like the gradients
            hh, ww = a["opt_f4"].shape[:2]
            ys, xs = np.mgrid[0:hh, 0:ww]
            idx = (((ys >> 2) * ((ww + 3) // 4) + (xs >> 2)) << 4) + ((ys & 3) << 2) + (xs & 3)
            q = a["opt"][idx]                                   # (h, w, 2)
            assert np.array_equal(q[..., 0], a["opt_f4"][..., 2].view(np.uint32))
            gx = (q[..., 1] & 0xffff).astype(np.uint16).view(np.int16).astype(np.float32) / 32764.0
            gy = (q[..., 1] >> 16).astype(np.uint16).view(np.int16).astype(np.float32) / 32764.0
            assert np.abs(gx[1:-1] - np.clip(ow[1:-1, :, 0], -1, 1)).max() <= 0.5 / 32764 + 1e-7
            assert np.abs(gy[1:-1] - np.clip(ow[1:-1, :, 1], -1, 1)).max() <= 0.5 / 32764 + 1e-7


def oracle_pyramid(orc, cam, n_levels, bgr, depth, n_percentage=0.3):
    from oracle import oracle as O

    cfg = O.PyrCfg(n_levels=n_levels)
    cfg.n_percentage = n_percentage
    p = O.build_pyramid(orc, cfg, cam, bgr, depth)
    O.make_keyframe(orc, p)
    return p


@pytest.mark.parametrize("seed,w,h,n_levels", [(4, 192, 144, 3), (6, 160, 120, 4)])
def test_pyramid_kernels_on_host_bit_exact(pyr_emu, orc32, seed, w, h, n_levels):
    p = synth_pair(seed, w, h)
    bgr, depth = p["key"]
    depth = depth.copy()
    depth[10:20, 30:60] = np.nan                               # holes for the hole-aware depth pyramid
    depth[h // 2, :] = 0.0
    po = oracle_pyramid(orc32, p["cam"], n_levels, bgr, depth)
    compare(run_pyramid(pyr_emu, bgr, depth, p["cam"], n_levels), po, n_levels)


def test_pyramid_kernels_on_host_fill_in_and_group_compaction(pyr_emu, orc32):
    """A high nPercentage forces fillInEdges on levels 1 and 2 (imgpyramidrgbd.cpp:188-196); eight frames per call take the
    group compaction of launch_compact (k_group_mask / k_group_scatter), the path of every batched build."""
    p = synth_pair(11, 96, 72)
    bgr, depth = p["key"]
    po = oracle_pyramid(orc32, p["cam"], 3, bgr, depth, n_percentage=1.1)
    assert all(not np.array_equal(po.edges[l], po.edges_orig[l]) for l in (1, 2))          # the fill-in changed both levels
    compare(run_pyramid(pyr_emu, bgr, depth, p["cam"], 3, n_percentage=1.1), po, 3)
    p = synth_pair(5, 64, 48)
    bgr, depth = p["key"]
    po = oracle_pyramid(orc32, p["cam"], 2, bgr, depth)
    compare(run_pyramid(pyr_emu, bgr, depth, p["cam"], 2, n_frames=8), po, 2)


def test_quality_vote_kernels_on_host(pyr_emu, orc32):
    """k_quality_scatter / k_quality_hist (TrackerNew::assessTrackingQuality, tracker.cpp:118-201) on the emulation layer vs the
    numpy restatement of the vote: exact histogram / overlap counts for three past frames."""
    from oracle import oracle as O
    from revo_b200 import synth

    w, h, lvl = 320, 240, 2
    s = synth.make_stream(31, 4, w, h, max_trans=0.02, max_rot_deg=1.0)
    cam = synth.intrinsics(w, h)
    cfg = O.PyrCfg(n_levels=3)
    pyrs = [O.build_pyramid(orc32, cfg, cam, *s["frames"][i]) for i in range(4)]
    T_w = [np.linalg.inv(s["T_w_c"][0]) @ s["T_w_c"][i] for i in range(4)]
    cur, est = pyrs[3], T_w[3].astype(np.float32)
    past_pts = [np.ascontiguousarray(pyrs[i].edges3d[lvl], np.float32) for i in range(3)]
    past_poses = [T_w[i].astype(np.float32) for i in range(3)]
    c = cur.cams[lvl]
    want = O.assess_tracking_quality(past_pts, past_poses, est, (c.fx, c.fy, c.cx, c.cy, c.w, c.h), cur.depth[lvl], cur.edges_orig[lvl])
    # inv(estimatedPose) * pastWorldPose in double, then float (revo_track_quality, capi.cu)
    rel = [np.linalg.inv(est.astype(np.float64)) @ P.astype(np.float64) for P in past_poses]
    R9 = np.ascontiguousarray(np.stack([r[:3, :3].T.reshape(-1) for r in rel]), np.float32)
    T3 = np.ascontiguousarray(np.stack([r[:3, 3] for r in rel]), np.float32)
    n_pts = np.ascontiguousarray([len(p) for p in past_pts], np.int32)
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in past_pts])
    depth = np.ascontiguousarray(cur.depth[lvl], np.float32)
    edges = np.ascontiguousarray(cur.edges_orig[lvl], np.uint8)
    counters = np.zeros(16, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    f = C.c_float
    rc = pyr_emu.emu_quality(C.c_int(3), ptrs, vp(n_pts), vp(R9), vp(T3), f(c.fx), f(c.fy), f(c.cx), f(c.cy), C.c_int(c.w), C.c_int(c.h),
                             vp(depth), vp(edges), f(0.1), f(5.2), vp(counters))
    assert rc == 0
    assert list(counters[:4]) == list(want["histogram"]) and list(counters[4:8]) == list(want["overlaps"]), (counters[:9], want)
    assert sum(counters[:4]) > 1000 and counters[5:8].sum() > 0


def test_colored_point_cloud_kernels_on_host(pyr_emu, orc32):
    """k_pyrdown_color + k_pcl_col_count / k_col_scan / k_pcl_col_scatter (ImgPyramidRGBD::generateColoredPcl,
    imgpyramidrgbd.cpp:279-327) on the emulation layer vs the loop restatement (pinned to the compiled reference in
    tests/test_oracle_ref.py): levels 0..2, edge cloud and dense cloud, 3 and 4 colour channels, bit-exact."""
    import cv2

    from oracle import oracle as O

    p = synth_pair(4, 160, 120)
    bgr, depth = p["key"]
    po = oracle_pyramid(orc32, p["cam"], 3, bgr, depth)
    bgra = np.ascontiguousarray(np.concatenate([bgr, np.full(bgr.shape[:2] + (1,), 255, np.uint8)], axis=2))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    f = C.c_float
    rgb = bgr
    for lvl in range(3):
        if lvl:
            rgb = cv2.pyrDown(rgb)
        c = po.cams[lvl]
        d = np.ascontiguousarray(po.depth[lvl], np.float32)
        e = np.ascontiguousarray(po.edges[lvl], np.uint8)
        for dense in (0, 1):
            want = O.generate_colored_pcl(rgb, d, e, (c.fx, c.fy, c.cx, c.cy, c.w, c.h), 0.1, 5.2, bool(dense))
            for img in (bgr, bgra):
                out = np.zeros((c.w * c.h, 8), np.float32)
                n = C.c_int(0)
                rc = pyr_emu.emu_colored_pcl(vp(np.ascontiguousarray(img)), C.c_int(img.shape[2]), C.c_int(160), C.c_int(120), C.c_int(lvl), vp(d),
                                             vp(e), C.c_int(c.w), C.c_int(c.h), f(c.fx), f(c.fy), f(c.cx), f(c.cy), C.c_int(dense), f(0.1),
                                             f(5.2), vp(out), C.c_int(c.w * c.h), C.byref(n))
                assert rc == 0 and n.value == want.shape[1], (lvl, dense, n.value, want.shape)
                assert np.array_equal(out[:n.value].T, want), (lvl, dense)
        assert want.shape[1] > 100


def test_distance_transform_of_sparse_and_empty_edge_maps_on_host(pyr_emu, orc32):
    """k_edt_cols / k_edt_rows where the search span matters: a flat image (no edge at all: every distance is OpenCV's 65536)
    and an image whose only edges sit in a narrow vertical band (most columns have no edge: the row search is clamped to the
    span of columns that have one) -- bit-exact against the oracle's cv2-pinned distance transform."""
    p = synth_pair(7, 160, 120)
    _, depth = p["key"]
    flat = np.full((120, 160, 3), 90, np.uint8)
    po = oracle_pyramid(orc32, p["cam"], 2, flat, depth)
    assert not po.edges[0].any() and float(po.dt[0].min()) == 65536.0
    compare(run_pyramid(pyr_emu, flat, depth, p["cam"], 2), po, 2)
    band = flat.copy()
    band[20:100, 70:74] = 230                                   # one bright bar: edges in columns ~69..74 only
    po = oracle_pyramid(orc32, p["cam"], 2, band, depth)
    cols = np.nonzero(po.edges[0].any(axis=0))[0]
    assert 0 < cols.size < 12 and float(po.dt[0].max()) > 60.0
    compare(run_pyramid(pyr_emu, band, depth, p["cam"], 2), po, 2)
