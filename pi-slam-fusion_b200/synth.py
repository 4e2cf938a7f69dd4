"""Deterministic synthetic nadir survey sequences: camera, serpentine poses, frame content.

Pure numpy; no dependency on the oracle or the CUDA library.  Poses are the reference's 7-double stream order
`x y z qx qy qz qw` (GSLAM/core/SE3.h:105-117), camera-to-world; the ground plane is the identity pose.
"""
import math

import numpy as np

IDENTITY_POSE = np.array([0, 0, 0, 0, 0, 0, 1], np.float64)


def camera(w, h):
    """PinHoleParameters w h fx fy cx cy (Map2D.h:37-43): fx = fy = 0.9 W, principal point at the centre."""
    return np.array([w, h, 0.9 * w, 0.9 * w, w / 2.0, h / 2.0], np.float64)


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _qaxis(axis, ang):
    s = math.sin(ang / 2)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, math.cos(ang / 2)])


def frames_per_line(n, w, h, along=0.2, cross=0.4):
    """Line length that makes the surveyed area roughly square."""
    return max(1, min(n, int(round(math.sqrt(n * cross * w / (along * h))))))


def serpentine_poses(n, w, h, altitude=100.0, along=0.2, cross=0.4, jitter=False, seed=0, fpl=None,
                     subpixel=True):
    """n camera-to-world poses on a serpentine grid: along-track step `along`*H*GSD (world +-y), cross-track
    `cross`*W*GSD (world +x).  Nadir attitude = 180 deg about X, i.e. quaternion (1,0,0,0).  `subpixel`
    adds a fixed irrational-ish offset per frame so that no two frames are integer-pixel translates of each
    other (exact weight ties other than 0>=0 then have measure zero).  `jitter`: yaw U(-5,5) deg, roll/pitch
    N(0,1 deg), altitude +-2 %."""
    cam = camera(w, h)
    gsd = altitude / cam[2]
    rng = np.random.default_rng(seed)
    fpl = fpl or frames_per_line(n, w, h, along, cross)
    poses = np.zeros((n, 7), np.float64)
    q0 = np.array([1.0, 0.0, 0.0, 0.0])
    for k in range(n):
        line, i = divmod(k, fpl)
        j = i if line % 2 == 0 else fpl - 1 - i
        x = line * cross * w * gsd
        y = j * along * h * gsd
        z = altitude
        q = q0
        if subpixel:
            x += gsd * ((k * 0.6180339887498949) % 1.0)
            y += gsd * ((k * 0.7548776662466927) % 1.0)
        if jitter:
            yaw = math.radians(rng.uniform(-5, 5))
            roll = math.radians(rng.normal(0, 1))
            pitch = math.radians(rng.normal(0, 1))
            z *= 1 + rng.uniform(-0.02, 0.02)
            q = _qmul(_qmul(_qaxis((0, 0, 1), yaw), _qmul(_qaxis((0, 1, 0), pitch), _qaxis((1, 0, 0), roll))), q0)
            q = q / np.linalg.norm(q)
        poses[k] = [x, y, z, q[0], q[1], q[2], q[3]]
    return poses


def ground_texture(size=2048, seed=0):
    """size x size x 3 u8: low-pass noise (sigma 3 px) + linear gradient + 8-px checker.  Periodic use."""
    rng = np.random.default_rng(seed)
    noise = rng.random((size, size, 3), dtype=np.float32)
    try:
        from scipy.ndimage import gaussian_filter
        noise = gaussian_filter(noise, sigma=(3, 3, 0), mode="wrap")
    except Exception:  # pragma: no cover - scipy is in the image; box blur keeps the generator usable anyway
        k = 7
        pad = np.pad(noise, ((k, k), (k, k), (0, 0)), mode="wrap")
        cs = pad.cumsum(0).cumsum(1)
        noise = (cs[2 * k:, 2 * k:] - cs[:-2 * k, 2 * k:] - cs[2 * k:, :-2 * k] + cs[:-2 * k, :-2 * k]) / (4.0 * k * k)
    noise = (noise - noise.min()) / (noise.max() - noise.min() + 1e-9)
    yy, xx = np.mgrid[0:size, 0:size]
    grad = ((xx + yy) % size) / float(size)
    checker = (((xx // 8) + (yy // 8)) % 2).astype(np.float32)
    img = 150.0 * noise + 60.0 * grad[..., None] + 40.0 * checker[..., None]
    img[..., 1] *= 0.9
    img[..., 2] = 255.0 - img[..., 2] * 0.8
    return np.clip(img, 0, 255).astype(np.uint8)


class Sequence:
    """A synthetic survey: camera, plane, poses and lazily generated frames (crops of a periodic texture)."""

    def __init__(self, n, w, h, seed=0, jitter=False, noise=False, texture_size=2048, along=0.2, cross=0.4,
                 altitude=100.0, fpl=None, prepare_frames=20):
        self.n, self.w, self.h, self.seed = n, w, h, seed
        self.camera = camera(w, h)
        self.plane = IDENTITY_POSE.copy()
        self.poses = serpentine_poses(n, w, h, altitude, along, cross, jitter, seed, fpl)
        self.prepare_poses = self.poses[:min(prepare_frames, n)]
        self.gsd = altitude / self.camera[2]
        self.noise = noise
        self._tex = None
        self._tsize = texture_size

    @property
    def texture(self):
        if self._tex is None:
            self._tex = ground_texture(self._tsize, self.seed)
        return self._tex

    def frame_origin(self, k):
        """Integer texture coordinates (row0, col0) of image pixel (v=0,u=0); rows then run DOWN the texture
        reversed (camera y looks along world -y)."""
        x, y = self.poses[k, 0], self.poses[k, 1]
        col0 = int(round(x / self.gsd - self.camera[4]))
        row0 = int(round(y / self.gsd + self.camera[5]))
        return row0, col0

    def frame(self, k):
        if self.noise:  # parity stress case: i.i.d. uniform noise
            return np.random.default_rng(self.seed * 100003 + k).integers(0, 256, (self.h, self.w, 3), dtype=np.uint8)
        row0, col0 = self.frame_origin(k)
        t = self.texture
        rows = (row0 - np.arange(self.h)) % t.shape[0]
        cols = (col0 + np.arange(self.w)) % t.shape[1]
        return np.ascontiguousarray(t[rows[:, None], cols[None, :]])

    def frames(self, idx=None):
        idx = range(self.n) if idx is None else idx
        return np.stack([self.frame(k) for k in idx])
