"""Pin of the Map2DRender (type 4) oracle against REAL OpenCV (cv2 4.13) -- CPU only.

Map2DRender.cpp:52-310 is an inline copy of cv::detail::MultiBandBlender; its CV_16S branch and its disabled `#else`
branch ARE stock OpenCV, so the oracle's blends 2 and 1 are compared bit for bit with cv2.detail_MultiBandBlender fed
with the same warped images, masks and corners.  Blend 0 (what the reference executes: `>=` selection per level) has no
OpenCV counterpart and is compared with the same loop rebuilt from cv2 primitives (copyMakeBorder, pyrDown, pyrUp,
subtract).  The two warps of renderFrames (:586-587) are pinned against cv2.warpPerspective directly.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as O  # noqa: E402
import pi_slam_fusion_b200.synth as synth  # noqa: E402
from tests.cv2_reference import warp_nearest_249  # noqa: E402

cv2.setNumThreads(1)
SRC4 = np.array([[0, 0], [128, 0], [0, 72], [128, 72]], np.float32)


@pytest.fixture(autouse=True)
def _reset_mode():
    O.set_f32_mode(0)
    O.set_threads(1)
    yield
    O.set_f32_mode(0)


def random_h(rng):
    dst = (SRC4 * rng.uniform(0.8, 1.6) + rng.uniform(0, 30, (1, 2)) + rng.normal(0, 4, (4, 2))).astype(np.float32)
    return cv2.getPerspectiveTransform(SRC4, dst)


def test_warp_u8c3_linear_reflect_bit_exact():
    rng = np.random.default_rng(0)
    for _ in range(8):
        M = random_h(rng)
        src = rng.integers(0, 256, (72, 128, 3), dtype=np.uint8)
        dsize = (int(rng.integers(100, 300)), int(rng.integers(60, 200)))
        ref = cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        assert np.array_equal(ref, O.warp_u8c3_reflect(src, M, dsize))


def test_warp_u8c1_nearest():
    """Same rule as the f32 nearest warp: equal to cv2 except the half-pixel border band that cv2 >= 4.12 zeroes."""
    rng = np.random.default_rng(1)
    for _ in range(6):
        M = random_h(rng)
        src = rng.integers(1, 256, (72, 128), dtype=np.uint8)
        dsize = (260, 170)
        a = cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_NEAREST)
        b = O.warp_u8c1_nearest(src, M, dsize)
        assert np.array_equal(b, warp_nearest_249(src, M, dsize))
        bad = a != b
        assert (a[bad] == 0).all() and bad.mean() < 0.02


def test_render_weight_image():
    for w, h in ((128, 72), (65, 33), (1280, 720)):
        xc, yc = np.float32(w * 0.5), np.float32(h * 0.5)
        inv = np.float32(1.0 / np.sqrt(xc * xc + yc * yc, dtype=np.float32))
        i = np.arange(h, dtype=np.float32)[:, None]
        j = np.arange(w, dtype=np.float32)[None, :]
        dis = (i - yc) * (i - yc) + (j - xc) * (j - xc)
        dis = np.float32(1) - np.sqrt(dis, dtype=np.float32) * inv
        ref = np.maximum((dis * dis * np.float32(254)).astype(np.uint8), 1)
        assert np.array_equal(ref, O.render_weight_image(w, h))


def _seq(n=7, w=160, h=96, seed=3, jitter=True):
    return synth.Sequence(n, w, h, seed=seed, jitter=jitter, fpl=3, prepare_frames=4)


def _render(blend, seq, bands=0, f32_mode=0):
    O.set_f32_mode(f32_mode)
    o = O.OracleMap2D(O.TYPE_RENDER, render_blend=blend, render_bands=bands)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    rc, res = o.render_frames(seq.frames(), seq.poses)
    assert rc == 0 and (res == 0).all()
    return o


@pytest.mark.parametrize("blend,wt", [(1, cv2.CV_32F), (2, cv2.CV_16S)])
@pytest.mark.parametrize("bands", [0, 2, 5])
def test_weighted_sum_blends_equal_cv2_multibandblender(blend, wt, bands):
    seq = _seq()
    o = _render(blend, seq, bands, f32_mode=1)  # cv2 4.x float association for the CV_32F weight pyramid
    res, mask, nb, _ = o.render_get()
    b = cv2.detail_MultiBandBlender(0, bands if bands else nb, wt)
    if not bands:  # the reference's rule for the band count, Map2DRender.cpp:707-715
        bw = np.float32(np.sqrt(np.float32(res.shape[0] * res.shape[1]))) * 5.0 / np.float32(100.0)
        assert nb == int(np.ceil(np.log(np.float32(bw)) / np.log(2.0)) - 1.0)
    b.prepare((0, 0, res.shape[1], res.shape[0]))
    assert b.numBands() == nb
    for i in range(seq.n):
        img, m, (cx, cy) = o.render_warped(i)
        b.feed(img, m, (cx, cy))
    ref, ref_mask = b.blend(None, None)
    assert np.array_equal(ref_mask, mask)
    assert np.array_equal(ref, res)
    assert mask.any() and not mask.all()


def _select_blend_cv2(imgs, masks, corners, W, H, nb):
    """Map2DRender.cpp:102-213 + :256-300 with the `#if 1` selection, from cv2 primitives only."""
    nb = min(nb, int(np.ceil(np.log(max(W, H)) / np.log(2.0))))
    dst_l, dst_w = [], []
    r, c = H, W
    for i in range(nb + 1):
        if i:
            r, c = (r + 1) // 2, (c + 1) // 2
        dst_l.append(np.zeros((r, c, 3), np.int16))
        dst_w.append(np.zeros((r, c), np.float32))
    for img, mask, (tlx, tly) in zip(imgs, masks, corners):
        ih, iw = mask.shape
        gap = 3 * (1 << nb)
        tnx, tny = max(0, tlx - gap), max(0, tly - gap)
        bnx, bny = min(W, tlx + iw + gap), min(H, tly + ih + gap)
        tnx, tny = (tnx >> nb) << nb, (tny >> nb) << nb
        width, height = bnx - tnx, bny - tny
        width += ((1 << nb) - width % (1 << nb)) % (1 << nb)
        height += ((1 << nb) - height % (1 << nb)) % (1 << nb)
        bnx, bny = tnx + width, tny + height
        dy, dx = max(bny - H, 0), max(bnx - W, 0)
        tnx, bnx, tny, bny = tnx - dx, bnx - dx, tny - dy, bny - dy
        top, left = tly - tny, tlx - tnx
        bottom, right = bny - tly - ih, bnx - tlx - iw
        bordered = cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_REFLECT)
        # createLaplacePyr, 8U branch
        lap = []
        cur, down = bordered, cv2.pyrDown(bordered)
        for i in range(1, nb):
            nxt = cv2.pyrDown(down)
            up = cv2.pyrUp(down, dstsize=(cur.shape[1], cur.shape[0]))
            lap.append(cv2.subtract(cur, up, dtype=cv2.CV_16S))
            cur, down = down, nxt
        up = cv2.pyrUp(down, dstsize=(cur.shape[1], cur.shape[0]))
        lap.append(cv2.subtract(cur, up, dtype=cv2.CV_16S))
        lap.append(down.astype(np.int16))
        wm = cv2.copyMakeBorder(mask.astype(np.float32) * np.float32(1.0 / 255.0), top, bottom, left, right, cv2.BORDER_CONSTANT, value=0)
        wp = [wm]
        for i in range(nb):
            wp.append(cv2.pyrDown(wp[-1]))
        x_tl, y_tl, x_br, y_br = tnx, tny, bnx, bny
        for i in range(nb + 1):
            dw = dst_w[i][y_tl:y_br, x_tl:x_br]
            dl = dst_l[i][y_tl:y_br, x_tl:x_br]
            take = wp[i] >= dw
            dw[take] = wp[i][take]
            dl[take] = lap[i][take]
            x_tl, y_tl, x_br, y_br = x_tl // 2, y_tl // 2, x_br // 2, y_br // 2
    for i in range(nb, 0, -1):
        up = cv2.pyrUp(dst_l[i], dstsize=(dst_l[i - 1].shape[1], dst_l[i - 1].shape[0]))
        dst_l[i - 1] = cv2.add(up, dst_l[i - 1])
    mask = (dst_w[0] > np.float32(1e-5)).astype(np.uint8) * 255
    res = dst_l[0].copy()
    res[mask == 0] = 0
    return res, mask, nb


@pytest.mark.parametrize("blend,wt", [(1, cv2.CV_32F), (2, cv2.CV_16S)])
def test_weighted_sum_blends_on_noise_and_exactly_nadir_frames(blend, wt):
    """i.i.d. noise content (every rounding counts) on exactly nadir poses (affine homographies), larger frames: still the
    real cv2.detail_MultiBandBlender bit for bit."""
    seq = synth.Sequence(9, 320, 180, seed=11, jitter=False, noise=True, fpl=3, prepare_frames=5)
    o = _render(blend, seq, 0, f32_mode=1)
    res, mask, nb, _ = o.render_get()
    b = cv2.detail_MultiBandBlender(0, nb, wt)
    b.prepare((0, 0, res.shape[1], res.shape[0]))
    for i in range(seq.n):
        img, m, c = o.render_warped(i)
        b.feed(img, m, c)
    ref, ref_mask = b.blend(None, None)
    assert np.array_equal(ref_mask, mask) and np.array_equal(ref, res) and mask.any()


@pytest.mark.parametrize("bands", [0, 3])
def test_selection_blend_equals_cv2_primitives(bands):
    seq = _seq(n=8, seed=5)
    o = _render(0, seq, bands, f32_mode=1)
    res, mask, nb, _ = o.render_get()
    w = [o.render_warped(i) for i in range(seq.n)]
    ref, ref_mask, ref_nb = _select_blend_cv2([x[0] for x in w], [x[1] for x in w], [x[2] for x in w], res.shape[1], res.shape[0], nb)
    assert ref_nb == nb
    assert np.array_equal(ref_mask, mask)
    assert np.array_equal(ref, res)


def test_render_warps_and_corners_equal_cv2():
    """The per-frame half of renderFrames (:536-603, :640-651) restated with numpy + cv2: sizes, corners, warped image (bit
    exact) and mask (2.4.9 nearest rule)."""
    seq = _seq(n=6, seed=9)
    o = _render(0, seq)
    g = o.grid()
    mn, lp = g["min"], g["length_pixel"]
    lpi = 1.0 / lp
    cam = seq.camera
    wimg = O.render_weight_image(seq.w, seq.h)
    frames = seq.frames()
    _, _, _, (tx0, ty0) = o.render_get()
    ele = 256 * lp
    minx, miny = mn[0] + ele * tx0, mn[1] + ele * ty0   # no spreadMap in this sequence: absolute == grid coordinates
    for i in range(seq.n):
        x, y, z, qx, qy, qz, qw = seq.poses[i]
        R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                      [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                      [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
        pts = []
        for (u, v) in ((0, 0), (cam[0], 0), (0, cam[1]), (cam[0], cam[1])):
            a = R @ np.array([(u - cam[4]) / cam[2], (v - cam[5]) / cam[3], 1.0])
            s = z / a[2]
            pts.append((x - a[0] * s, y - a[1] * s))
        pts = np.array(pts)
        cmin, cmax = pts.min(0), pts.max(0)
        img, mask, (cx, cy) = o.render_warped(i)
        size = (int((cmax[0] - cmin[0]) * lpi), int((cmax[1] - cmin[1]) * lpi))
        assert abs(mask.shape[1] - size[0]) <= 1 and abs(mask.shape[0] - size[1]) <= 1   # numpy's rotation differs in the last ulp
        size = (mask.shape[1], mask.shape[0])
        assert abs(cx - int((np.float32(cmin[0]) - minx) * lpi)) <= 1 and abs(cy - int((np.float32(cmin[1]) - miny) * lpi)) <= 1
        src = np.array([[0, 0], [cam[0], 0], [0, cam[1]], [cam[0], cam[1]]], np.float32)
        dst = ((pts - cmin) * lpi).astype(np.float32)
        M = cv2.getPerspectiveTransform(src, dst)
        ref = cv2.warpPerspective(frames[i], M, size, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        # the oracle's pose algebra (quaternion sandwich) and numpy's rotation matrix agree to ~1e-13: same px except where a
        # 1/32-px coordinate rounds the other way
        assert (ref != img).mean() < 2e-3
        assert (warp_nearest_249(wimg, M, size) != mask).mean() < 2e-3


def test_render_is_rejected_through_feed_and_needs_prepare():
    seq = _seq(n=3)
    o = O.OracleMap2D(O.TYPE_RENDER)
    rc, _ = o.render_frames(seq.frames(), seq.poses)
    assert rc != 0
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert not o.feed(seq.frame(0), seq.poses[0])   # Map2DRender::renderFrame returns false (thread = false)
    assert o.render_get() is None


def test_render_batch_of_oblique_frames_only_gives_a_blank_canvas():
    """Every frame fails the 0.4 test (Map2DRender.cpp:548-557) -> min/max stay at the origin (:532) -> the one tile around the
    origin is 'blended' from nothing: renderFrames still returns true, with an all-masked canvas."""
    seq = _seq(n=4)
    poses = seq.poses.copy()
    for k in range(seq.n):
        poses[k, 3:] = synth._qmul(synth._qaxis((0, 1, 0), np.radians(75.0)), np.array([1.0, 0.0, 0.0, 0.0]))
    o = O.OracleMap2D(O.TYPE_RENDER)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    rc, res = o.render_frames(seq.frames(), poses)
    assert rc == 0 and (res == 1).all()
    r16, mask, nb, _ = o.render_get()
    assert r16.shape == (256, 256, 3) and not r16.any() and not mask.any() and nb == 3
