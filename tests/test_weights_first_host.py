"""Host-side proofs behind the weights-first multi-band pipeline (kernels.cu "WEIGHTS-FIRST variant"), on the CPU:
  * the reach table (which cells of which Gaussian level a winner depends on) against a brute-force walk of the actual
    pyrUp / pyrDown taps, borders included;
  * the FP32 weight-coordinate fast path of mbw_warp_kernel: wherever it does NOT flag a px as ambiguous, its rounded
    source coordinate equals the exact FP64 one OpenCV computes."""
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


def fma32(a, b, c):
    """__fmaf_rn for float32 arrays: the product of two floats is exact in float64, one rounding to float32 at the end
    (the intermediate float64 sum rounds first; a double rounding can differ from the true FMA by 1 ulp in ~1e-9 of cases,
    far inside the 48-ulp band this test is about)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def lean_path(M, xs, ys, sw, sh):
    """numpy restatement of the EXPERIMENTAL mbw_weights4_lean FP32 pass: explicit FMAs, one reciprocal per 4-px run carried
    by Newton steps, rounding by the 1.5*2^23 trick."""
    f = np.float32
    mf = M.astype(np.float32)
    xf, yf = xs.astype(np.float32), ys.astype(np.float32)
    x4 = (np.floor(xs / 4) * 4).astype(np.float32)
    jj = (xs - np.floor(xs / 4) * 4).astype(int)
    den0, nx0, ny0 = fma32(mf[7], yf, mf[8]), fma32(mf[1], yf, mf[2]), fma32(mf[4], yf, mf[5])
    wa, wb = fma32(mf[6], x4, den0), fma32(mf[6], x4 + f(3), den0)
    wmin = np.minimum(wa, wb)
    ok = wmin > f(1e-3)
    newton = abs(mf[6]) * f(3) < f(1e-4) * wmin
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        r = f(1) / wa
        r3 = r.copy()
        for j in range(1, 4):
            den = fma32(mf[6], x4 + f(j), den0)
            step = np.where(newton, fma32(r3, fma32(-den, r3, f(1)), r3), f(1) / den).astype(np.float32)
            r3 = step
            r = np.where(jj >= j, step, r).astype(np.float32)     # the reciprocal this px ends up with
        fx, fy = fma32(mf[0], xf, nx0) * r, fma32(mf[3], xf, ny0) * r
        rmax = np.maximum(f(1) / wa, r3)
        magx = fma32(abs(mf[0]), x4 + f(3), fma32(abs(mf[1]), yf, abs(mf[2]))) * rmax
        magy = fma32(abs(mf[3]), x4 + f(3), fma32(abs(mf[4]), yf, abs(mf[5]))) * rmax
        thr_x, thr_y = fma32(f(48) * f(5.97e-8), magx, f(1e-6)), fma32(f(48) * f(5.97e-8), magy, f(1e-6))
        magic = f(12582912.0)
        tx, ty = (fx + magic).astype(np.float32), (fy + magic).astype(np.float32)
        rx, ry = tx - magic, ty - magic
        ix = tx.view(np.int32) - 0x4B400000
        iy = ty.view(np.int32) - 0x4B400000
        sane = (abs(fx) < f(2097152)) & (abs(fy) < f(2097152))
        slack_x, slack_y = f(4) * f(5.97e-8) * magx, f(4) * f(5.97e-8) * magy     # __fdividef vs IEEE division
        amb = ~ok | (sane & ((f(0.5) - abs(fx - rx) < thr_x - slack_x) | (f(0.5) - abs(fy - ry) < thr_y - slack_y)))
    assert np.array_equal(ix[sane & ok].astype(np.float32), rx[sane & ok]), "bits(t) - 0x4B400000 must be the rounded value"
    far = ok & ~sane      # millions of px away: the kernel writes weight 0 without asking the exact path
    return np.where(far, f(-1e9), rx), np.where(far, f(-1e9), ry), amb, far


@pytest.mark.parametrize("variant", ["default", "lean"])
@pytest.mark.parametrize("w,h,n,tilt", [(1280, 720, 120, False), (4000, 3000, 40, False), (1920, 1080, 60, True), (320, 180, 40, True)])
def test_fp32_weight_coordinates_agree_with_fp64_when_not_ambiguous(w, h, n, tilt, variant):
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
        if variant == "lean":
            rx, ry, amb, far = lean_path(M, xs, ys, w, h)
            inside = (ex >= 0) & (ex < w) & (ey >= 0) & (ey < h)
            assert not (far & inside).any(), "a px declared 'far outside' must really be outside the frame"
            sure = ~amb & ~far
        else:
            rx, ry, amb = fast_path(M, xs, ys, w, h)
            sure = ~amb
        assert np.array_equal(rx[sure], ex[sure]) and np.array_equal(ry[sure], ey[sure]), "frame %d" % k
        checked += int(sure.sum())
        ambiguous += int(amb.sum())
    assert checked > 100000
    assert ambiguous < 0.25 * (checked + ambiguous), "the fast path must decide the great majority of px"


def level_weight_upper_bound(mf, x0, y0, x1, y1, sw, sh, weight_type):
    """Upper bound of a frame's level-l weight over a set of level-l px, given the level-0 rect [x0,x1] x [y0,y1]
    (region px, inclusive) that contains the supports of those px.  FP32 throughout, as a kernel would evaluate it:
    the image of the rect under the inverse homography is the convex hull of its 4 mapped corners (denominators
    positive), its bounding box grown by 1 px covers every rounded sampling position, and the weight image decreases
    with the distance to the frame centre — so the weight at the box's point nearest to the centre bounds them all.
    PyrDown is a convex combination (borders reflect inwards), so coarser levels inherit the bound; (1 + 3e-5) covers
    the float rounding of up to five filter passes.  A plan for round 2 (cull entries in the decide stage without
    loading them); validated here against real weight pyramids."""
    f = np.float32
    xs = np.array([x0, x1, x0, x1], np.float32)
    ys = np.array([y0, y0, y1, y1], np.float32)
    den = mf[6] * xs + mf[7] * ys + mf[8]
    if (den <= f(1e-3)).any():
        return np.inf
    sx, sy = (mf[0] * xs + mf[1] * ys + mf[2]) / den, (mf[3] * xs + mf[4] * ys + mf[5]) / den
    bx0, bx1, by0, by1 = sx.min() - f(1), sx.max() + f(1), sy.min() - f(1), sy.max() + f(1)
    if bx1 < f(-0.5) or bx0 > f(sw) - f(0.5) or by1 < f(-0.5) or by0 > f(sh) - f(0.5):
        return 0.0   # the whole set samples outside the frame
    xc, yc = f(sw // 2), f(sh // 2)
    dx = max(f(0), bx0 - xc, xc - bx1)
    dy = max(f(0), by0 - yc, yc - by1)
    dis = f(1) - min(np.sqrt(dx * dx + dy * dy) / np.sqrt(xc * xc + yc * yc), f(1))
    v = dis if weight_type == 0 else dis * dis
    return float(max(v, f(1e-5)) * f(1 + 3e-5) + f(1e-7))


@pytest.mark.parametrize("w,h,weight_type,tilt", [(320, 180, 0, False), (320, 180, 1, True), (1280, 720, 0, True)])
def test_level_weight_upper_bound_is_conservative(w, h, weight_type, tilt):
    seq = synth.Sequence(12, w, h, seed=17, jitter=True)
    poses = seq.poses.copy()
    if tilt:
        rng = np.random.default_rng(2)
        for k in range(seq.n):
            q = synth._qmul(synth._qmul(synth._qaxis((0, 0, 1), rng.uniform(-3, 3)),
                                        synth._qmul(synth._qaxis((0, 1, 0), rng.uniform(-0.4, 0.4)), synth._qaxis((1, 0, 0), rng.uniform(-0.4, 0.4)))),
                            np.array([1.0, 0, 0, 0]))
            poses[k, 3:] = q / np.linalg.norm(q)
    m = O.OracleMap2D.create(3, weight_type=weight_type)
    assert m.prepare(seq.plane, seq.camera, poses[:6])
    rects, hinv = m.compute_bounds(poses)
    wimg = O.weight_image_f32(w, h, weight_type)
    rng = np.random.default_rng(3)
    levels, checked, useful = 6, 0, 0
    for k in range(0, seq.n, 3 if w > 1000 else 1):
        if rects[k, 2] <= rects[k, 0]:
            continue
        nx, ny = (rects[k, 2] - rects[k, 0]) * 256, (rects[k, 3] - rects[k, 1]) * 256
        pyr = [O.warp_f32_nearest(wimg, np.linalg.inv(hinv[k]), (nx, ny))]
        for _ in range(levels - 1):
            pyr.append(O.pyrdown_f32(pyr[-1]))
        mf = hinv[k].reshape(9).astype(np.float32)
        for l in range(levels):
            R = 0 if l == 0 else (2 << l) - 2          # level-0 half-support of a level-l px: 2^(l+1) - 2
            H, Wd = pyr[l].shape
            for _ in range(150):
                cw, chh = min(64, Wd), min(max(2, 2 * (32 * 2 // max(min(64, Wd), 1))), H)   # a warp's px set: 64 x 2, or full rows
                X0, Y0 = int(rng.integers(0, Wd - cw + 1)), int(rng.integers(0, H - chh + 1)) & ~1
                actual = float(pyr[l][Y0:Y0 + chh, X0:X0 + cw].max())
                ub = level_weight_upper_bound(mf, (X0 << l) - R, (Y0 << l) - R, ((X0 + cw - 1) << l) + R, ((Y0 + chh - 1) << l) + R,
                                              w, h, weight_type)
                assert ub >= actual, (k, l, X0, Y0, ub, actual)
                checked += 1
                useful += ub < 0.5
    assert checked > 1000 and useful > 0.1 * checked, "the bound must also be tight enough to reject far-away frames"


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
