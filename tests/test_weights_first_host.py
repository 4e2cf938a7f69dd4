"""Host-side proofs behind the weights-first multi-band pipeline (csrc/kernels_wf.cu), on the CPU:
  * the reach tables (which cells of which Gaussian / weight level a winner / a competitive cell depends on) against a
    brute-force walk of the actual pyrUp / pyrDown taps, borders included;
  * the FP32 weight-coordinate fast path of mbw_warp_kernel: wherever it does NOT flag a px as ambiguous, its rounded
    source coordinate equals the exact FP64 one OpenCV computes;
  * the closed-form cell bounds the pipeline culls with (csrc/bounds.h, called through m2d_cell_weight_bounds -- the very
    code the bounds kernel runs) against real weight pyramids built by the oracle;
  * the order-free form of the reference's sequential `>=` rule."""
import numpy as np
import pytest

import pi_slam_fusion_b200.map2d as m2d
import pi_slam_fusion_b200.synth as synth
from oracle import oracle as O


def reflect101(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * n - 2 - p
    return p


@pytest.mark.parametrize("levels", [1, 2, 3, 4, 5, 6])
def test_reach_table_covers_every_tap(levels):
    lo, hi = (t.astype(int) for t in m2d.reach_table(levels))
    tiles = 3                                     # a 3-tile-wide region: borders on both sides are exercised
    size = [tiles * (256 >> l) for l in range(levels)]
    worst = np.zeros((6, 6, 2), int)
    for m in range(levels):
        for p in range(size[m]):                  # every winner position of level m (1-D: the filters are separable)
            need = {m: {p}}
            if m + 1 < levels:                    # lap_quad: pyrUp taps around the quad's coarse px
                n1 = size[m + 1]
                i = (p & ~1) >> 1
                need[m + 1] = {(1 if n1 > 1 else 0) if i - 1 < 0 else i - 1, i, min(i + 1, n1 - 1)}
            top = max(need)
            for k in range(top, 0, -1):           # pyrDown taps, BORDER_REFLECT_101
                below = need.setdefault(k - 1, set())
                for u in need[k]:
                    below.update(reflect101(2 * u + d - 2, size[k - 1]) for d in range(5))
            cwin = (p << m) >> 5
            for k, pts in need.items():
                assert lo[m][k] != 255, (m, k)
                cells = [(q << k) >> 5 for q in pts]
                assert min(cells) >= cwin - lo[m][k] and max(cells) <= cwin + hi[m][k], (levels, m, p, k, min(cells), max(cells), cwin)
                worst[m][k] = np.maximum(worst[m][k], [cwin - min(cells), max(cells) - cwin])
    for m in range(levels):                       # ... and the table is tight: every entry is reached by some winner
        for k in range(levels):
            if lo[m][k] != 255:
                assert (worst[m][k] == [lo[m][k], hi[m][k]]).all(), (levels, m, k, worst[m][k], lo[m][k], hi[m][k])
            else:
                assert k > m + 1


def fast_path(M, xs, ys, sw, sh):
    """numpy float32 restatement of the FP32 pass of mbw_warp_kernel (one px at a time; the kernel shares den0/nx0/ny0
    per 4-px run, which is the same arithmetic).  Returns rounded coords and the 'ambiguous' flag."""
    f = np.float32
    mf = M.astype(np.float32)
    xf, yf = xs.astype(np.float32), ys.astype(np.float32)
    x4 = (np.floor(xs / 4) * 4).astype(np.float32)           # first px of the thread's run
    den0 = mf[7] * yf + mf[8]
    nx0, ny0 = mf[1] * yf + mf[2], mf[4] * yf + mf[5]
    wa, wb = mf[6] * x4 + den0, mf[6] * (x4 + f(3)) + den0
    ok = (wa > f(1e-3)) & (wb > f(1e-3))
    with np.errstate(divide="ignore", invalid="ignore"):
        rmax = np.maximum(f(1) / wa, f(1) / wb)
        magx = (abs(mf[0]) * (x4 + f(3)) + abs(mf[1]) * yf + abs(mf[2])) * rmax
        magy = (abs(mf[3]) * (x4 + f(3)) + abs(mf[4]) * yf + abs(mf[5])) * rmax
        thr_x, thr_y = f(48) * f(5.97e-8) * magx + f(1e-6), f(48) * f(5.97e-8) * magy + f(1e-6)
        r = f(1) / (mf[6] * xf + den0)
        fx, fy = (mf[0] * xf + nx0) * r, (mf[3] * xf + ny0) * r
    rx, ry = np.rint(fx), np.rint(fy)
    # __fdividef is within 2 ulp of the IEEE quotient used here: widen the emulated ambiguity band accordingly
    slack_x, slack_y = f(4) * f(5.97e-8) * magx, f(4) * f(5.97e-8) * magy
    amb = ~ok | (f(0.5) - abs(fx - rx) < thr_x - slack_x) | (f(0.5) - abs(fy - ry) < thr_y - slack_y)
    return rx, ry, amb


@pytest.mark.parametrize("w,h,n,tilt", [(1280, 720, 120, False), (4000, 3000, 40, False), (1920, 1080, 60, True), (320, 180, 40, True)])
def test_fp32_weight_coordinates_agree_with_fp64_when_not_ambiguous(w, h, n, tilt):
    seq = synth.Sequence(n, w, h, seed=11, jitter=True)
    poses = seq.poses.copy()
    if tilt:   # strong roll/pitch (still accepted by the ray.down >= 0.4 test): large perspective terms
        rng = np.random.default_rng(1)
        for k in range(n):
            q = synth._qmul(synth._qmul(synth._qaxis((0, 0, 1), rng.uniform(-3, 3)),
                                        synth._qmul(synth._qaxis((0, 1, 0), rng.uniform(-0.5, 0.5)), synth._qaxis((1, 0, 0), rng.uniform(-0.5, 0.5)))),
                            np.array([1.0, 0, 0, 0]))
            poses[k, 3:] = q / np.linalg.norm(q)
    m = O.OracleMap2D.create(3)
    assert m.prepare(seq.plane, seq.camera, poses[:20])
    rects, hinv = m.compute_bounds(poses)
    rng = np.random.default_rng(0)
    checked = ambiguous = 0
    for k in range(n):
        if rects[k, 0] == -1 and rects[k, 2] == -1:
            continue
        M = hinv[k].reshape(9)
        nx, ny = (rects[k, 2] - rects[k, 0]) * 256, (rects[k, 3] - rects[k, 1]) * 256
        xs = rng.integers(0, nx, 60000).astype(np.float64)
        ys = rng.integers(0, ny, 60000).astype(np.float64)
        xb = np.floor(xs / 64) * 64                  # OpenCV's association: X0 = M0*xb + M1*y + M2, then + M0*x1
        x1 = xs - xb
        W = 1.0 / (M[6] * xb + M[7] * ys + M[8] + M[6] * x1)
        ex = np.rint((M[0] * xb + M[1] * ys + M[2] + M[0] * x1) * W)
        ey = np.rint((M[3] * xb + M[4] * ys + M[5] + M[3] * x1) * W)
        rx, ry, amb = fast_path(M, xs, ys, w, h)
        sure = ~amb
        assert np.array_equal(rx[sure], ex[sure]) and np.array_equal(ry[sure], ey[sure]), "frame %d" % k
        checked += int(sure.sum())
        ambiguous += int(amb.sum())
    assert checked > 100000
    assert ambiguous < 0.25 * (checked + ambiguous), "the fast path must decide the great majority of px"


@pytest.mark.parametrize("levels", [1, 2, 3, 4, 5, 6])
def test_weight_reach_table_covers_every_tap(levels):
    lo, hi = (t.astype(int) for t in m2d.weight_reach_table(levels))
    tiles = 3
    size = [tiles * (256 >> l) for l in range(levels)]
    worst = np.zeros((6, 6, 2), int)
    for m in range(levels):
        for p in range(size[m]):                  # every px of weight level m that a competitive cell may contain
            need = {m: {p}}
            for k in range(m, 0, -1):             # pyrDown taps, BORDER_REFLECT_101
                need[k - 1] = {reflect101(2 * u + d - 2, size[k - 1]) for u in need[k] for d in range(5)}
            c = (p << m) >> 5
            for k, pts in need.items():
                assert lo[m][k] != 255
                cells = [(q << k) >> 5 for q in pts]
                assert min(cells) >= c - lo[m][k] and max(cells) <= c + hi[m][k], (levels, m, p, k)
                worst[m][k] = np.maximum(worst[m][k], [c - min(cells), max(cells) - c])
    for m in range(levels):
        for k in range(6):
            if k <= m:
                assert (worst[m][k] == [lo[m][k], hi[m][k]]).all(), (levels, m, k, worst[m][k], lo[m][k], hi[m][k])
            else:
                assert lo[m][k] == 255 and hi[m][k] == 255


def tilted_poses(seq, seed, yaw=3.0, tilt=0.4):
    rng = np.random.default_rng(seed)
    poses = seq.poses.copy()
    for k in range(seq.n):
        q = synth._qmul(synth._qmul(synth._qaxis((0, 0, 1), rng.uniform(-yaw, yaw)),
                                    synth._qmul(synth._qaxis((0, 1, 0), rng.uniform(-tilt, tilt)), synth._qaxis((1, 0, 0), rng.uniform(-tilt, tilt)))),
                        np.array([1.0, 0, 0, 0]))
        poses[k, 3:] = q / np.linalg.norm(q)
    return poses


@pytest.mark.parametrize("w,h,weight_type,tilt,scale", [(320, 180, 0, False, 1.0), (320, 180, 1, True, 1.0), (1280, 720, 0, True, 1.0),
                                                     (1280, 720, 0, False, 1.0), (640, 360, 1, False, 1.7), (640, 360, 0, True, 0.6)])
def test_cell_weight_bounds_are_conservative(w, h, weight_type, tilt, scale):
    """lo <= W_l(u) <= hi for EVERY px of EVERY cell of every level of real weight pyramids (oracle primitives: nearest
    warp of the weight image, five f32 pyrDowns with the region's reflect-101 border), nadir / jittered / strongly tilted
    frames, both weight types, map scales below and above 1.  Also: the bounds are tight enough to be useful."""
    seq = synth.Sequence(12, w, h, seed=17, jitter=True)
    poses = tilted_poses(seq, 2) if tilt else seq.poses
    m = O.OracleMap2D.create(3, weight_type=weight_type, scale=scale)
    assert m.prepare(seq.plane, seq.camera, poses[:6])
    rects, hinv = m.compute_bounds(poses)
    wimg = O.weight_image_f32(w, h, weight_type)
    levels, cells, tight, gap = 6, 0, 0, []
    for k in range(0, seq.n, 4 if w > 1000 else 2):
        if rects[k, 2] <= rects[k, 0]:
            continue
        nx, ny = int(rects[k, 2] - rects[k, 0]), int(rects[k, 3] - rects[k, 1])
        pyr = [O.warp_f32_nearest(wimg, np.linalg.inv(hinv[k]), (nx * 256, ny * 256))]
        for _ in range(levels - 1):
            pyr.append(O.pyrdown_f32(pyr[-1]))
        for l in range(levels):
            B = max(32 >> l, 1)
            blk = pyr[l].reshape(ny * 8, B, nx * 8, B)
            mn, mx = blk.min(axis=(1, 3)), blk.max(axis=(1, 3))
            for cy in range(ny * 8):
                for cx in range(nx * 8):
                    lo, hi = m2d.cell_weight_bounds(hinv[k], nx, ny, w, h, weight_type, l, cx, cy)
                    assert lo <= mn[cy, cx] and mx[cy, cx] <= hi, (k, l, cx, cy, lo, float(mn[cy, cx]), float(mx[cy, cx]), hi)
                    cells += 1
                    if lo > 0:
                        tight += 1
                        gap.append(hi - lo)
    print("cells %d tight %.3f mean gap %.3f" % (cells, tight / cells, float(np.mean(gap))))
    assert cells > 5000 and tight > (0.04 if w * scale < 600 else 0.15) * cells, (cells, tight)   # a 320x180 frame is 10 x 5.6 cells in a 24 x 16-cell region
    assert np.mean(gap) < (0.45 if w * scale < 600 else 0.1), "interior cells: a band of a few percent of the weight range (a cell's 32 px are a large part of a small frame)"


def test_latest_maximum_rule_equals_the_sequential_scan():
    """The reference updates a px with `if (srcW >= dstW)` frame after frame (MultiBandMap2DCPU.cpp:539-547).  What that
    leaves behind is order-free: the largest weight wins, the LATEST frame among equals, and the state that was there
    before only survives if it is strictly larger than every frame.  (Basis of the experimental best-first decide stage.)"""
    rng = np.random.default_rng(4)
    for _ in range(2000):
        n = int(rng.integers(1, 9))
        vals = rng.choice([0.0, 1e-5, 0.25, 0.5, 0.75], size=n + 1).astype(np.float32)   # few distinct values: many ties
        fresh = bool(rng.integers(0, 2))
        state, frames = (-np.inf if fresh else vals[0]), vals[1:]
        bw, best = state, -1
        for k, s in enumerate(frames):            # sequential scan in feed order
            if s >= bw:
                bw, best = s, k
        order = rng.permutation(n)                # any visiting order
        bw2, best2 = state, -1
        for k in order:
            s = frames[k]
            if s > bw2 or (s == bw2 and k > best2):
                bw2, best2 = s, int(k)
        assert (bw, best) == (bw2, best2)


def _warp_coords(Mi, width, height):
    """The 1/32-px integer coordinates cv::warpPerspectiveInvoker computes for a width x height destination (the oracle's
    warp_row_coords: 64-px column blocks, X0 = M0*xb + M1*y + M2 advanced by M0*x1), vectorised in float64."""
    x = np.arange(width)
    xb = ((x // 64) * 64).astype(np.float64)[None, :]
    x1 = (x % 64).astype(np.float64)[None, :]
    y = np.arange(height, dtype=np.float64)[:, None]
    X0 = (Mi[0, 0] * xb + Mi[0, 1] * y) + Mi[0, 2]
    Y0 = (Mi[1, 0] * xb + Mi[1, 1] * y) + Mi[1, 2]
    W0 = (Mi[2, 0] * xb + Mi[2, 1] * y) + Mi[2, 2]
    W = W0 + Mi[2, 0] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        W = np.where(W != 0, 32.0 / W, 0.0)
    fX = np.clip((X0 + Mi[0, 0] * x1) * W, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + Mi[1, 0] * x1) * W, -2147483648.0, 2147483647.0)
    return np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)


def _reflect(p, n):
    """cv::borderInterpolate(BORDER_REFLECT), vectorised (any number of folds)."""
    p = np.mod(p, 2 * n)
    return np.where(p >= n, 2 * n - 1 - p, p)


@pytest.mark.parametrize("w,h,tilt,scale", [(640, 360, False, 1.0), (640, 360, True, 1.0), (1280, 720, True, 1.0), (320, 180, True, 0.5),
                                            (640, 360, True, 2.2), (96, 64, True, 1.0)])
def test_pull_rect_contains_every_tap(w, h, tilt, scale):
    """Pull mode (kernels_wf.cu mbs_mark / mbs_pull) copies, for every needed 32 x 32 cell, the source rectangle
    m2d_pull_cell_rect returns.  Walk the bilinear taps of EVERY px of EVERY cell of the frames' regions exactly as the warp
    computes them (1/32-px rounding, saturate_cast<short>, BORDER_REFLECT) and check that none falls outside -- nadir, jittered
    and strongly tilted frames, map scales 0.5 - 2.2, a frame smaller than its reflection margin.  Also: the rectangle is not
    wastefully large (what is fetched stays close to what is read)."""
    seq = synth.Sequence(10, w, h, seed=23, jitter=True)
    poses = tilted_poses(seq, 2) if tilt else seq.poses
    m = O.OracleMap2D.create(3, scale=scale)
    assert m.prepare(seq.plane, seq.camera, poses[:5])
    rects, hinv = m.compute_bounds(poses)
    cells = waste_num = waste_den = 0
    for k in range(seq.n):
        if rects[k, 2] <= rects[k, 0]:
            continue
        nx, ny = int(rects[k, 2] - rects[k, 0]), int(rects[k, 3] - rects[k, 1])
        X, Y = _warp_coords(hinv[k], nx * 256, ny * 256)
        sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
        tx0, tx1, ty0, ty1 = _reflect(sx, w), _reflect(sx + 1, w), _reflect(sy, h), _reflect(sy + 1, h)
        txlo, txhi = np.minimum(tx0, tx1), np.maximum(tx0, tx1)
        tylo, tyhi = np.minimum(ty0, ty1), np.maximum(ty0, ty1)
        blk = lambda a, f: f(f(a.reshape(ny * 8, 32, nx * 8, 32), axis=3), axis=1)   # noqa: E731
        cxlo, cxhi, cylo, cyhi = blk(txlo, np.min), blk(txhi, np.max), blk(tylo, np.min), blk(tyhi, np.max)
        for cy in range(ny * 8):
            for cx in range(nx * 8):
                ok, (lox, hix, loy, hiy) = m2d.pull_cell_rect(hinv[k], cx * 32, cy * 32, w, h)
                assert 0 <= lox <= hix < w and 0 <= loy <= hiy < h
                assert lox <= cxlo[cy, cx] and cxhi[cy, cx] <= hix and loy <= cylo[cy, cx] and cyhi[cy, cx] <= hiy, \
                    (k, cx, cy, (lox, hix, loy, hiy), (int(cxlo[cy, cx]), int(cxhi[cy, cx]), int(cylo[cy, cx]), int(cyhi[cy, cx])))
                cells += 1
                inside = (sx[cy * 32:cy * 32 + 32, cx * 32:cx * 32 + 32] >= 0).all() and (sx[cy * 32:cy * 32 + 32, cx * 32:cx * 32 + 32] < w - 1).all() and \
                    (sy[cy * 32:cy * 32 + 32, cx * 32:cx * 32 + 32] >= 0).all() and (sy[cy * 32:cy * 32 + 32, cx * 32:cx * 32 + 32] < h - 1).all()
                if ok and inside:   # interior cells: fetched area vs the taps' own bounding box
                    waste_num += (hix - lox + 1) * (hiy - loy + 1)
                    waste_den += (cxhi[cy, cx] - cxlo[cy, cx] + 1) * (cyhi[cy, cx] - cylo[cy, cx] + 1)
    assert cells > 500
    if waste_den:
        assert waste_num / waste_den < (1.6 if scale <= 1.0 else 2.6), waste_num / waste_den   # 5 px of margin per axis on a 32/scale-px box
