"""Tier-2 pin: the two renderFrame recipes restated in Python on top of the REAL OpenCV primitives
(cv2 4.13: getPerspectiveTransform, warpPerspective, pyrDown, detail.createLaplacePyr,
detail.restoreImageFromLaplacePyr).  Used only to pin the C++ oracle (tests/test_oracle_vs_cv2.py) and to
generate tests/golden/*.  Geometry is an independent float64 restatement of Map2DCPU.cpp:44-92,163-233.
"""
import math

import numpy as np

try:
    import cv2
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None

ELE = 256


def qmul(a, b):  # SO3.h:435-442, (x,y,z,w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return (aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz)


def qinv(q):
    return (-q[0], -q[1], -q[2], q[3])


def qrot(q, p):  # SO3.h:445-450
    r = qmul(qmul(q, (p[0], p[1], p[2], 0.0)), qinv(q))
    return (r[0], r[1], r[2])


def pose_inv(p):  # SE3.h:70-73 ; p = (t, q)
    ri = qinv(p[1])
    v = qrot(ri, p[0])
    return ((-v[0], -v[1], -v[2]), ri)


def pose_mul(a, b):  # SE3.h:84-89
    rt = qrot(a[1], b[0])
    return ((a[0][0] + rt[0], a[0][1] + rt[1], a[0][2] + rt[2]), qmul(a[1], b[1]))


def pose7(v):
    v = [float(x) for x in v]
    return ((v[0], v[1], v[2]), (v[3], v[4], v[5], v[6]))


def warp_nearest_249(src, M, dsize):
    """warpPerspective(INTER_NEAREST, BORDER_CONSTANT 0) with OpenCV 2.4.9 semantics, in numpy float64.

    cv2 >= 4.12 ships a rewritten nearest kernel that tests the UNROUNDED coordinate against [0, n-1], so it
    zeroes the half-pixel bands (-0.5,0) and (n-1,n-0.5] that 2.4.9 (round first, then `(unsigned)sx < n`)
    keeps.  Everywhere else the two agree bit-exactly (tests/test_oracle_vs_cv2.py::test_warp_nearest)."""
    dw, dh = dsize
    Mi = cv2.invert(np.asarray(M, np.float64))[1]
    bh0 = min(16, dh)
    bw0 = min(1024 // bh0, dw)
    x = np.arange(dw)
    xb = ((x // bw0) * bw0).astype(np.float64)[None, :]
    x1 = (x % bw0).astype(np.float64)[None, :]
    y = np.arange(dh, dtype=np.float64)[:, None]
    X0 = (Mi[0, 0] * xb + Mi[0, 1] * y) + Mi[0, 2]
    Y0 = (Mi[1, 0] * xb + Mi[1, 1] * y) + Mi[1, 2]
    W0 = (Mi[2, 0] * xb + Mi[2, 1] * y) + Mi[2, 2]
    W = W0 + Mi[2, 0] * x1
    with np.errstate(divide="ignore"):
        W = np.where(W != 0, 1.0 / W, 0.0)
    fX = np.clip((X0 + Mi[0, 0] * x1) * W, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + Mi[1, 0] * x1) * W, -2147483648.0, 2147483647.0)
    sx = np.clip(np.rint(fX).astype(np.int64), -32768, 32767)
    sy = np.clip(np.rint(fY).astype(np.int64), -32768, 32767)
    inside = (sx >= 0) & (sx < src.shape[1]) & (sy >= 0) & (sy < src.shape[0])
    out = np.zeros((dh, dw), src.dtype)
    out[inside] = src[sy[inside], sx[inside]]
    return out


class Cv2Map2D:
    def __init__(self, type_, scale=1.0, resolution=0.0, weight_type=0, band_number=5, background=0):
        self.type = 1 if type_ == 2 else type_
        self.scale, self.resolution, self.weight_type, self.background = scale, resolution, weight_type, background
        self.bands = min(band_number, int(math.ceil(math.log(ELE) / math.log(2.0))))
        self.valid = False
        self.wimg = None

    def unproject(self, u, v):
        return ((u - self.cx) * self.fxinv, (v - self.cy) * self.fyinv, 1.0)

    def prepare(self, plane, cam, poses):
        if len(poses) == 0 or cam[0] <= 0 or cam[1] <= 0 or cam[2] == 0 or cam[3] == 0:
            return False
        self.cw, self.ch, self.fx, self.fy, self.cx, self.cy = [float(c) for c in cam]
        self.fxinv, self.fyinv = 1.0 / self.fx, 1.0 / self.fy
        self.plane = pose7(plane)
        pinv = pose_inv(self.plane)
        mx, mn = [-1e10] * 3, [1e10] * 3
        for p in poses:
            t = pose_mul(pinv, pose7(p))[0]
            for i in range(3):
                mx[i] = max(mx[i], t[i])
                mn[i] = min(mn[i], t[i])
        if mn[2] * mx[2] <= 0:
            return False
        if self.type == 3:
            hgt = mx[2] if mx[2] > 0 else -mn[2]
        else:
            hgt = mn[2] if mn[2] > 0 else -mx[2]
        a, b = self.unproject(self.cw, self.ch), self.unproject(0.0, 0.0)
        lx, ly = a[0] - b[0], a[1] - b[1]
        radius = 0.5 * hgt * math.sqrt(lx * lx + ly * ly)
        lp = self.resolution if self.type == 3 else 0.0
        if not lp:
            lp = 2 * radius / math.sqrt(self.cw * self.cw + self.ch * self.ch)
            lp /= self.scale
        self.lp, self.lpinv = lp, 1.0 / lp
        mn = [mn[0] - radius, mn[1] - radius, mn[2]]
        mx = [mx[0] + radius, mx[1] + radius, mx[2]]
        c = [0.5 * (mn[i] + mx[i]) for i in range(3)]
        mn = [2 * mn[i] - c[i] for i in range(3)]
        mx = [2 * mx[i] - c[i] for i in range(3)]
        self.ele = ELE * lp
        self.eleinv = 1.0 / self.ele
        self.w = int(math.ceil((mx[0] - mn[0]) / self.ele))
        self.h = int(math.ceil((mx[1] - mn[1]) / self.ele))
        mx[0] = mn[0] + self.ele * self.w
        mx[1] = mn[1] + self.ele * self.h
        self.min, self.max = mn, mx
        self.tiles = {}
        self.valid = True
        return True

    def spread(self, xmin, ymin, xmax, ymax):
        x0 = min(int(math.floor((xmin - self.min[0]) * self.eleinv)), 0)
        y0 = min(int(math.floor((ymin - self.min[1]) * self.eleinv)), 0)
        x1 = max(int(math.ceil((xmax - self.min[0]) * self.eleinv)), self.w)
        y1 = max(int(math.ceil((ymax - self.min[1]) * self.eleinv)), self.h)
        nw, nh = x1 - x0, y1 - y0
        nminx = self.min[0] + self.ele * x0
        nminy = self.min[1] + self.ele * y0
        self.tiles = {(x - x0, y - y0): t for (x, y), t in self.tiles.items()}
        self.min = [nminx, nminy, self.min[2]]
        self.max = [nminx + nw * self.ele, nminy + nh * self.ele, self.max[2]]
        self.w, self.h = nw, nh

    def bounds(self, pose):
        f = pose_mul(pose_inv(self.plane), pose7(pose))
        t = f[0]
        ipts = [(0.0, 0.0), (self.cw, 0.0), (0.0, self.ch), (self.cw, self.ch)]
        down = (0.0, 0.0, 1.0) if t[2] < 0 else (0.0, 0.0, -1.0)
        pts = []
        for (u, v) in ipts:
            ax = qrot(f[1], self.unproject(u, v))
            if ax[0] * down[0] + ax[1] * down[1] + ax[2] * down[2] < 0.4:
                return None
            s = t[2] / ax[2]
            pts.append((t[0] - ax[0] * s, t[1] - ax[1] * s))
        return ipts, pts

    def feed(self, img, pose):
        if not self.valid or img.shape[1] != self.cw or img.shape[0] != self.ch:
            return False
        b = self.bounds(pose)
        if b is None:
            return False
        ipts, pts = b
        xmin = min(p[0] for p in pts); xmax = max(p[0] for p in pts)
        ymin = min(p[1] for p in pts); ymax = max(p[1] for p in pts)
        if xmin < self.min[0] or xmax > self.max[0] or ymin < self.min[1] or ymax > self.max[1]:
            self.spread(xmin, ymin, xmax, ymax)
        x0 = int(math.floor((xmin - self.min[0]) * self.eleinv)); y0 = int(math.floor((ymin - self.min[1]) * self.eleinv))
        x1 = int(math.ceil((xmax - self.min[0]) * self.eleinv)); y1 = int(math.ceil((ymax - self.min[1]) * self.eleinv))
        if x0 < 0 or y0 < 0 or x1 > self.w or y1 > self.h or x0 >= x1 or y0 >= y1:
            return False
        ox = self.min[0] + self.ele * x0
        oy = self.min[1] + self.ele * y0
        src = np.array(ipts, np.float32)
        dst = np.array([((p[0] - ox) * self.lpinv, (p[1] - oy) * self.lpinv) for p in pts], np.float64).astype(np.float32)
        M = cv2.getPerspectiveTransform(src, dst)
        self.last_rect = (x0, y0, x1, y1)
        self.last_M = M
        dsize = ((x1 - x0) * ELE, (y1 - y0) * ELE)
        if self.type == 3:
            self._render_multiband(img, M, dsize, x0, y0, x1, y1)
        else:
            self._render_weighted(img, M, dsize, x0, y0, x1, y1)
        return True

    # Map2DCPU.cpp:236-258 / MultiBandMap2DCPU.cpp:396-418, in numpy float32 (same op order)
    def _weights(self, w, h):
        f = np.float32
        xc, yc = f(w // 2), f(h // 2)
        dmax = np.sqrt(xc * xc + yc * yc, dtype=np.float32)
        ii = (np.arange(h, dtype=np.int32).astype(np.float32) - yc)[:, None]
        jj = (np.arange(w, dtype=np.int32).astype(np.float32) - xc)[None, :]
        dis = (ii * ii + jj * jj).astype(np.float32)
        dis = (f(1) - np.sqrt(dis, dtype=np.float32) / dmax).astype(np.float32)
        return dis

    def _render_weighted(self, img, M, dsize, x0, y0, x1, y1):
        h, w = img.shape[:2]
        if self.wimg is None or self.wimg.shape != (h, w):
            dis = self._weights(w, h)
            if self.weight_type == 0:
                a = (dis.astype(np.float64) * 254.0).astype(np.uint8)
            else:
                a = ((dis * dis).astype(np.float32) * np.float32(254)).astype(np.float32).astype(np.uint8)
            self.wimg = np.maximum(a, 2)
        src = np.dstack([img, self.wimg])
        dst = cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_LINEAR)
        for x in range(x0, x1):
            for y in range(y0, y1):
                t = self.tiles.get((x, y))
                if t is None:
                    t = self.tiles[(x, y)] = np.zeros((ELE, ELE, 4), np.uint8)
                d = dst[(y - y0) * ELE:(y - y0 + 1) * ELE, (x - x0) * ELE:(x - x0 + 1) * ELE]
                m = t[..., 3] < d[..., 3]
                t[m] = d[m]

    def _render_multiband(self, img, M, dsize, x0, y0, x1, y1):
        h, w = img.shape[:2]
        if self.wimg is None or self.wimg.shape != (h, w):
            dis = self._weights(w, h)
            v = dis if self.weight_type == 0 else (dis * dis).astype(np.float32)
            v = v.copy()
            v[v.astype(np.float64) <= 1e-5] = np.float32(1e-5)
            self.wimg = v
        img16 = img.astype(np.int16)
        image_warped = cv2.warpPerspective(img16, M, dsize, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        weight_warped = warp_nearest_249(self.wimg, M, dsize)
        pyr = cv2.detail.createLaplacePyr(image_warped.copy(), self.bands, [cv2.UMat() for _ in range(self.bands + 1)])
        pyr = [p.get() for p in pyr]
        pw = [weight_warped]
        for i in range(self.bands):
            pw.append(cv2.pyrDown(pw[i]))
        self.last_pyr, self.last_pw = pyr, pw
        for x in range(x0, x1):
            for y in range(y0, y1):
                t = self.tiles.get((x, y))
                fresh = t is None
                if fresh:
                    t = self.tiles[(x, y)] = ([None] * (self.bands + 1), [None] * (self.bands + 1))
                n = ELE
                for i in range(self.bands + 1):
                    sl = (slice((y - y0) * n, (y - y0 + 1) * n), slice((x - x0) * n, (x - x0 + 1) * n))
                    if fresh:
                        t[0][i] = pyr[i][sl].copy()
                        t[1][i] = pw[i][sl].copy()
                    else:
                        m = pw[i][sl] >= t[1][i]
                        t[0][i][m] = pyr[i][sl][m]
                        t[1][i][m] = pw[i][sl][m]
                    n //= 2

    def get_tile(self, tx, ty, level=0):
        t = self.tiles.get((tx, ty))
        if t is None:
            return None
        return t if self.type != 3 else (t[0][level], t[1][level])

    def get_image(self):
        if not self.tiles:
            return None
        xs = [k[0] for k in self.tiles]; ys = [k[1] for k in self.tiles]
        x0, x1, y0, y1 = min(xs), max(xs) + 1, min(ys), max(ys) + 1
        tw, th = x1 - x0, y1 - y0
        if self.type != 3:
            out = np.zeros((th * ELE, tw * ELE, 4), np.uint8)
            for (x, y), t in self.tiles.items():
                out[(y - y0) * ELE:(y - y0 + 1) * ELE, (x - x0) * ELE:(x - x0 + 1) * ELE] = t
            return out, (x0, y0)
        pyr = [np.zeros((th * (ELE >> i), tw * (ELE >> i), 3), np.int16) for i in range(self.bands + 1)]
        w0 = np.zeros((th * ELE, tw * ELE), np.float32)
        for (x, y), t in self.tiles.items():
            for i in range(self.bands + 1):
                n = ELE >> i
                pyr[i][(y - y0) * n:(y - y0 + 1) * n, (x - x0) * n:(x - x0 + 1) * n] = t[0][i]
            w0[(y - y0) * ELE:(y - y0 + 1) * ELE, (x - x0) * ELE:(x - x0 + 1) * ELE] = t[1][0]
        res = cv2.detail.restoreImageFromLaplacePyr([cv2.UMat(p) for p in pyr])
        r0 = res[0].get()
        out = np.clip(r0, 0, 255).astype(np.uint8)
        out[w0 == 0] = self.background
        return out, (x0, y0)
