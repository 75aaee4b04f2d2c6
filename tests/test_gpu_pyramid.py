"""GPU parity: ImgPyramidRGBD construction + makeKeyframe through the C ABI vs the oracle
(cv2 4.13 + C restatement of the reference loops). Byte/integer work: bit-exact."""
import numpy as np
import pytest

from conftest import synth_pair

pytestmark = pytest.mark.gpu


def _settings(cam, n_levels):
    from revo_b200 import api

    fx, fy, cx, cy, w, h = cam
    return api.ImgPyramidSettings(PYR_MIN_LVL=n_levels - 1, PYR_MAX_LVL=0, width=w, height=h, fx=fx, fy=fy, cx=cx, cy=cy)


def _oracle_pyr(orc, cam, n_levels, bgr, depth, keyframe=True, n_percentage=0.3):
    from oracle import oracle as O

    cfg = O.PyrCfg(n_levels=n_levels, n_percentage=n_percentage)
    p = O.build_pyramid(orc, cfg, cam, bgr, depth, backend="cv2")
    if keyframe:
        O.make_keyframe(orc, p, backend="cv2")
    return p


def _compare(pg, po, n_levels, keyframe=True):
    for l in range(n_levels):
        assert np.array_equal(pg.returnGray(l), po.gray[l]), f"gray L{l}"
        assert np.array_equal(pg.returnDepth(l), po.depth[l], equal_nan=True), f"depth L{l}"
        assert np.array_equal(pg._download(l, 3, np.uint8, lambda c: (c.height, c.width)), po.edges_orig[l]), f"canny L{l}"
        assert np.array_equal(pg.returnEdges(l), po.edges[l]), f"edges L{l}"
        assert np.array_equal(pg.returnHist(l), po.hist[l]), f"hist L{l}"
        e3 = pg.return3DEdges(l)
        assert e3.shape == po.edges3d[l].shape, f"edge count L{l}: {e3.shape} vs {po.edges3d[l].shape}"
        assert np.array_equal(e3, po.edges3d[l]), f"edges3d L{l}"
        assert pg.returnNumEdges(l) == len(po.edges3d[l])
        # the tracker's tile-major list is a permutation of the reference list
        dev = pg.return3DEdgesDeviceOrder(l)
        a = dev[np.lexsort(dev.T[::-1])]
        b = po.edges3d[l][np.lexsort(po.edges3d[l].T[::-1])]
        assert np.array_equal(a, b), f"device-order list L{l}"
        if keyframe:
            assert np.array_equal(pg.returnDistTransform(l), po.dt[l]), f"dt L{l}"
            assert np.array_equal(pg.returnOptimizationStructure(l), po.opt[l]), f"opt L{l}"


@pytest.mark.parametrize("seed,w,h,n_levels", [(1, 640, 480, 3), (2, 640, 480, 4), (3, 320, 240, 3), (5, 1920, 1080, 3)])
def test_pyramid_bit_exact(ctx, orc32, seed, w, h, n_levels):
    from revo_b200 import api

    p = synth_pair(seed, w, h)
    bgr, depth = p["key"]
    st = _settings(p["cam"], n_levels)
    pg = api.ImgPyramidRGBD(ctx, st, None, bgr, depth, 1.5)
    pg.makeKeyframe()
    po = _oracle_pyr(orc32, p["cam"], n_levels, bgr, depth)
    _compare(pg, po, n_levels)
    assert pg.returnTimestamp() == 1.5


def test_pyramid_fill_in_path(ctx, orc32):
    """Sparse texture -> fraction of non-empty patches < nPercentage -> fillInEdges runs (imgpyramidrgbd.cpp:188-196)."""
    from revo_b200 import api

    rng = np.random.default_rng(7)
    h, w = 480, 640
    bgr = np.full((h, w, 3), 90, np.uint8)
    # a few strong thin structures only: most 20x20 patches stay empty
    for k in range(6):
        x0, y0 = int(rng.integers(20, w - 120)), int(rng.integers(20, h - 120))
        bgr[y0:y0 + int(rng.integers(31, 99)), x0:x0 + int(rng.integers(31, 99))] = int(rng.integers(150, 255))
    bgr = np.clip(bgr.astype(np.int16) + rng.integers(-2, 3, bgr.shape), 0, 255).astype(np.uint8)
    depth = (1.0 + 0.5 * rng.random((h, w))).astype(np.float32)
    depth[rng.random((h, w)) < 0.05] = 0.0
    depth[rng.random((h, w)) < 0.01] = np.nan
    cam = (525.0, 525.0, 319.5, 239.5, w, h)
    st = _settings(cam, 3)
    pg = api.ImgPyramidRGBD(ctx, st, None, bgr, depth)
    pg.makeKeyframe()
    po = _oracle_pyr(orc32, cam, 3, bgr, depth)
    assert any(po.filled), "test scene must trigger the fill-in path"
    _compare(pg, po, 3)


def test_pyramid_noise_and_empty(ctx, orc32):
    """Pure noise (max edge density, many tiny components) and a constant image (no edges: DT = 65536 like cv2's own trueDistTrans)."""
    from revo_b200 import api

    rng = np.random.default_rng(3)
    h, w = 240, 320
    cam = (260.0, 260.0, 159.5, 119.5, w, h)
    st = _settings(cam, 3)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    depth = np.full((h, w), 2.0, np.float32)
    for bgr in (noise, np.full((h, w, 3), 77, np.uint8)):
        pg = api.ImgPyramidRGBD(ctx, st, None, bgr, depth)
        pg.makeKeyframe()
        po = _oracle_pyr(orc32, cam, 3, bgr, depth)
        _compare(pg, po, 3)


def test_pyramid_batch_matches_single(ctx, orc32):
    from revo_b200 import api

    ps = [synth_pair(s, 320, 240) for s in (11, 12, 13)]
    st = _settings(ps[0]["cam"], 3)
    bgr = np.stack([p["cur"][0] for p in ps])
    depth = np.stack([p["cur"][1] for p in ps])
    pyrs = api.ImgPyramidRGBD.create_batch(ctx, st, bgr, depth, timestamps=[0.1, 0.2, 0.3])
    api.ImgPyramidRGBD.makeKeyframes(ctx, pyrs)
    for i, pg in enumerate(pyrs):
        po = _oracle_pyr(orc32, ps[i]["cam"], 3, bgr[i], depth[i])
        _compare(pg, po, 3)
        assert abs(pg.returnTimestamp() - 0.1 * (i + 1)) < 1e-12


def test_bgra_input_and_errors(ctx, orc32):
    from revo_b200 import api

    p = synth_pair(3, 320, 240)
    bgr, depth = p["key"]
    bgra = np.concatenate([bgr, np.full(bgr.shape[:2] + (1,), 255, np.uint8)], axis=2)
    st = _settings(p["cam"], 3)
    pg = api.ImgPyramidRGBD(ctx, st, None, bgra, depth)
    po = _oracle_pyr(orc32, p["cam"], 3, bgr, depth, keyframe=False)
    _compare(pg, po, 3, keyframe=False)
    # returnOptimizationStructure before makeKeyframe: error code instead of the reference's exit(0)
    with pytest.raises(api.RevoError) as ei:
        pg.returnOptimizationStructure(0)
    assert ei.value.code == 4
    with pytest.raises(api.RevoError) as ei:
        pg.returnGray(5)
    assert ei.value.code == 6


def test_uint16_depth_wire_format(ctx, orc32):
    """16-bit raw depth (TUM wire format) converted on the device == float(raw) * (1.0f / 5000) with one rounding per
    pixel, i.e. cv::Mat::convertTo(CV_32FC1, 1.0f / DEPTH_SCALE_FACTOR) of the reference's reader
    (io/iowrapperRGBD.cpp:327; checked here against cv2.multiply); the pyramids are then identical."""
    import cv2

    from revo_b200 import api

    ps = [synth_pair(s) for s in (1, 2)]
    bgr = np.stack([p["key"][0] for p in ps])
    raw = np.stack([np.round(p["key"][1].astype(np.float64) * 5000.0).astype(np.uint16) for p in ps])
    raw[0, :7, :13] = 65535
    st = _settings(ps[0]["cam"], 3)
    scale = np.float32(1.0) / np.float32(5000.0)
    depth_host = raw.astype(np.float32) * scale
    assert np.array_equal(depth_host[0], cv2.multiply(raw[0].astype(np.float32), float(scale)))
    b16 = api.PyramidBatch(ctx, st, bgr, raw, 2, depth_scale_factor=5000.0)
    b32 = api.PyramidBatch(ctx, st, bgr, depth_host, 2)
    ctx.synchronize()
    for i in range(2):
        for l in range(3):
            assert np.array_equal(b16[i].returnDepth(l), b32[i].returnDepth(l)), (i, l)
            assert np.array_equal(b16[i].returnEdges(l), b32[i].returnEdges(l))
            assert np.array_equal(b16[i].return3DEdges(l), b32[i].return3DEdges(l))
    assert np.array_equal(b16[0].returnDepth(0), depth_host[0])
    b16.destroy()
    b32.destroy()


def test_canny_tile_fallback_path(ctx, orc32):
    """Widths that are not a multiple of 4 take the tile / union-find Canny (TMA tile load when the width allows it, plain
    loads otherwise) instead of the bit-mask pipeline: same bit-exact result (single-level pyramid: odd sizes allowed)."""
    from revo_b200 import api

    p = synth_pair(3, 640, 480)
    bgr, depth = p["key"]
    for w, h in ((322, 242), (250, 200)):
        b, d = np.ascontiguousarray(bgr[:h, :w]), np.ascontiguousarray(depth[:h, :w])
        cam = (300.0, 300.0, w / 2.0, h / 2.0, w, h)
        st = _settings(cam, 1)
        pg = api.ImgPyramidRGBD(ctx, st, None, b, d)
        pg.makeKeyframe()
        po = _oracle_pyr(orc32, cam, 1, b, d)
        _compare(pg, po, 1)


def test_colored_point_cloud_device_kernel(ctx, orc32):
    """revo_pyr_colored_pcl (ImgPyramidRGBD::generateColoredPcl, viewer export, imgpyramidrgbd.cpp:279-327): colour pyrDown +
    count / scan / scatter on the device against the loop restatement over the oracle pyramid, levels 0..2, edge and dense
    clouds, 3- and 4-channel colour images; the count-only call and the capacity check of the C ABI."""
    import ctypes as C

    import cv2

    from oracle import oracle as O
    from revo_b200 import api

    for seed, (w, h) in ((4, (320, 240)), (9, (640, 480))):
        p = synth_pair(seed, w, h)
        bgr, depth = p["key"]
        st = _settings(p["cam"], 3)
        pg = api.ImgPyramidRGBD(ctx, st, None, bgr, depth)
        po = _oracle_pyr(orc32, p["cam"], 3, bgr, depth, keyframe=False)
        bgra = np.concatenate([bgr, np.full(bgr.shape[:2] + (1,), 255, np.uint8)], axis=2)
        rgb = bgr
        for lvl in range(3):
            if lvl:
                rgb = cv2.pyrDown(rgb)
            c = po.cams[lvl]
            for dense in (False, True):
                want = O.generate_colored_pcl(rgb, po.depth[lvl], po.edges[lvl], (c.fx, c.fy, c.cx, c.cy, c.w, c.h), 0.1, 5.2, dense)
                got = pg.generateColoredPcl(lvl, dense)
                assert got.shape == want.shape and np.array_equal(got, want), (seed, lvl, dense)
                assert np.array_equal(pg.generateColoredPcl(lvl, dense, rgb=bgra), want)
            # the edge cloud is the 3-D edge list with colours
            assert np.array_equal(pg.generateColoredPcl(lvl)[:4].T, pg.return3DEdges(lvl))
        assert pg.generateColoredPcl(3).shape == (8, 0)
        n = C.c_int(0)
        small = np.zeros((4, 8), np.float32)
        rc = ctx.lib.revo_pyr_colored_pcl(ctx.h, pg.h, 0, 1, bgr.ctypes.data, 3, small.ctypes.data, 4, C.byref(n))
        assert rc == api.REVO_ERR_BUFFER_TOO_SMALL and n.value == int(np.isfinite(po.depth[0]).sum() - (po.depth[0] <= 0.1).sum()
                                                                      - (po.depth[0] >= 5.2).sum())
