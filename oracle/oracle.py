"""ctypes binding of the CPU oracle (oracle/map2d_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The class mirrors the reference's Map2D interface (Map2DFusion/Map2D.h:79-98): create / prepare / feed /
save (as get_image), so parity tests read like calls into the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmap2d_oracle.so")
MAX_LEVELS = 9


class Config(C.Structure):
    """m2d_config, include/map2d_b200.h"""
    _fields_ = [("scale", C.c_double), ("resolution", C.c_double), ("weight_type", C.c_int),
                ("band_number", C.c_int), ("force_float", C.c_int), ("background", C.c_int),
                ("thread", C.c_int), ("device", C.c_int), ("shard_rank", C.c_int), ("shard_count", C.c_int),
                ("shard_axis", C.c_int), ("shard_span", C.c_int), ("collect_stats", C.c_int),
                ("batch_frames", C.c_int), ("f32_mode", C.c_int), ("render_blend", C.c_int),
                ("render_bands", C.c_int)]


class Stats(C.Structure):
    """m2d_stats, include/map2d_b200.h"""
    _fields_ = [("frames_fed", C.c_uint64), ("frames_fused", C.c_uint64), ("input_px", C.c_uint64),
                ("region_px", C.c_uint64 * MAX_LEVELS), ("fresh_px", C.c_uint64 * MAX_LEVELS),
                ("win_px", C.c_uint64 * MAX_LEVELS), ("footprint_px", C.c_uint64), ("need_px", C.c_uint64 * MAX_LEVELS),
                ("needw_px", C.c_uint64 * MAX_LEVELS)]

    def as_dict(self):
        return {"frames_fed": self.frames_fed, "frames_fused": self.frames_fused, "input_px": self.input_px,
                "region_px": list(self.region_px), "fresh_px": list(self.fresh_px), "win_px": list(self.win_px),
                "footprint_px": self.footprint_px, "need_px": list(self.need_px), "needw_px": list(self.needw_px)}


def default_config(**kw):
    c = Config(scale=1.0, resolution=0.0, weight_type=0, band_number=5, force_float=0, background=0, thread=0,
               device=0, shard_rank=0, shard_count=1, shard_axis=1, shard_span=4, collect_stats=0, batch_frames=0)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def build(force=False):
    src = os.path.join(_HERE, "map2d_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "map2d_b200.h")
    inl = os.path.join(_HERE, "render_oracle.inl")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(inl))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libmap2d_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, fp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.POINTER(Config), C.POINTER(vp)]
        L.orc_destroy.argtypes = [vp]
        L.orc_destroy.restype = None
        L.orc_prepare.argtypes = [vp, dp, dp, C.c_int, dp]
        L.orc_feed.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, dp]
        L.orc_feed_poses.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int)]
        L.orc_plan_rects.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int)]
        L.orc_set_shard.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_get_grid.argtypes = [vp, ip, ip, dp, dp, dp]
        L.orc_last_rect.argtypes = [vp, ip]
        L.orc_get_tile.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_get_image.argtypes = [vp, vp, ip, ip, ip, ip, ip]
        L.orc_get_tile_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, ip]
        L.orc_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.orc_tile_bytes.argtypes = [vp]
        L.orc_tile_bytes.restype = C.c_size_t
        L.orc_tile_count.argtypes = [vp]
        L.orc_export_tiles.argtypes = [vp, C.c_int, ip, vp, C.c_int, ip]
        L.orc_import_tiles.argtypes = [vp, C.c_int, ip, vp, C.c_int]
        L.orc_export_tiles_rect.argtypes = [vp, ip, C.c_int, ip, vp, C.c_int, ip]
        L.orc_drop_tiles_rect.argtypes = [vp, ip, ip]
        L.orc_tile_bbox.argtypes = [vp, ip]
        L.orc_get_image_rect.argtypes = [vp, vp, C.c_int, ip, ip, ip, ip, ip]
        L.orc_compute_bounds.argtypes = [vp, C.c_int, dp, ip, dp]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = None
        L.orc_set_f32_mode.argtypes = [C.c_int]
        L.orc_set_f32_mode.restype = None
        L.orc_get_perspective_transform.argtypes = [fp, fp, dp]
        L.orc_invert3x3.argtypes = [dp, dp]
        for n in ("orc_warp_u8c4", "orc_warp_s16c3_reflect", "orc_warp_f32_nearest"):
            getattr(L, n).argtypes = [vp, C.c_int, C.c_int, dp, vp, C.c_int, C.c_int]
        L.orc_pyrdown_s16.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.orc_pyrdown_s16.restype = None
        L.orc_pyrdown_f32.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_pyrdown_f32.restype = None
        L.orc_pyrup_s16.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.orc_pyrup_s16.restype = None
        L.orc_weight_image_u8.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.orc_weight_image_u8.restype = None
        L.orc_weight_image_f32.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.orc_weight_image_f32.restype = None
        for n in ("orc_warp_u8c3_reflect", "orc_warp_u8c1_nearest"):
            getattr(L, n).argtypes = [vp, C.c_int, C.c_int, dp, vp, C.c_int, C.c_int]
        L.orc_render_weight_image.argtypes = [C.c_int, C.c_int, vp]
        L.orc_render_weight_image.restype = None
        L.orc_render_frames.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_size_t, dp, ip]
        L.orc_render_get.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip]
        L.orc_render_warped.argtypes = [vp, C.c_int, vp, vp, ip, ip, ip, ip]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def set_threads(n):
    lib().orc_set_threads(int(n))


def set_f32_mode(mode):
    """0 = OpenCV 2.4.9 float association (the reference; default), 1 = OpenCV 4.x/cv2 4.13 (pin tests only)."""
    lib().orc_set_f32_mode(int(mode))


# ---------------------------------------------------------------- primitives (for pinning against cv2)
def get_perspective_transform(src, dst):
    s = np.ascontiguousarray(src, np.float32).reshape(8)
    d = np.ascontiguousarray(dst, np.float32).reshape(8)
    M = np.zeros(9, np.float64)
    rc = lib().orc_get_perspective_transform(s.ctypes.data_as(C.POINTER(C.c_float)),
                                             d.ctypes.data_as(C.POINTER(C.c_float)), _dptr(M))
    if rc:
        raise ValueError("singular")
    return M.reshape(3, 3)


def invert3x3(M):
    M = np.ascontiguousarray(M, np.float64).reshape(9)
    Mi = np.zeros(9, np.float64)
    if lib().orc_invert3x3(_dptr(M), _dptr(Mi)):
        raise ValueError("singular")
    return Mi.reshape(3, 3)


def _warp(fn, src, M, dsize, dtype, cn):
    src = np.ascontiguousarray(src, dtype)
    dw, dh = dsize
    dst = np.zeros((dh, dw, cn) if cn > 1 else (dh, dw), dtype)
    M = np.ascontiguousarray(M, np.float64).reshape(9)
    if fn(src.ctypes.data, src.shape[0], src.shape[1], _dptr(M), dst.ctypes.data, dh, dw):
        raise ValueError("singular")
    return dst


def warp_u8c4(src, M, dsize):
    return _warp(lib().orc_warp_u8c4, src, M, dsize, np.uint8, 4)


def warp_s16c3_reflect(src, M, dsize):
    return _warp(lib().orc_warp_s16c3_reflect, src, M, dsize, np.int16, 3)


def warp_f32_nearest(src, M, dsize):
    return _warp(lib().orc_warp_f32_nearest, src, M, dsize, np.float32, 1)


def pyrdown_s16(a):
    a = np.ascontiguousarray(a, np.int16)
    cn = 1 if a.ndim == 2 else a.shape[2]
    oshape = ((a.shape[0] + 1) // 2, (a.shape[1] + 1) // 2) + (() if a.ndim == 2 else (cn,))
    out = np.zeros(oshape, np.int16)
    lib().orc_pyrdown_s16(a.ctypes.data, a.shape[0], a.shape[1], cn, out.ctypes.data)
    return out


def pyrdown_f32(a):
    a = np.ascontiguousarray(a, np.float32)
    out = np.zeros(((a.shape[0] + 1) // 2, (a.shape[1] + 1) // 2), np.float32)
    lib().orc_pyrdown_f32(a.ctypes.data, a.shape[0], a.shape[1], out.ctypes.data)
    return out


def pyrup_s16(a):
    a = np.ascontiguousarray(a, np.int16)
    cn = 1 if a.ndim == 2 else a.shape[2]
    oshape = (a.shape[0] * 2, a.shape[1] * 2) + (() if a.ndim == 2 else (cn,))
    out = np.zeros(oshape, np.int16)
    lib().orc_pyrup_s16(a.ctypes.data, a.shape[0], a.shape[1], cn, out.ctypes.data)
    return out


def warp_u8c3_reflect(src, M, dsize):
    """cv2.warpPerspective(8UC3, INTER_LINEAR, BORDER_REFLECT) -- Map2DRender.cpp:586"""
    return _warp(lib().orc_warp_u8c3_reflect, src, M, dsize, np.uint8, 3)


def warp_u8c1_nearest(src, M, dsize):
    """cv2.warpPerspective(8UC1, INTER_NEAREST, BORDER_CONSTANT 0) -- Map2DRender.cpp:587"""
    return _warp(lib().orc_warp_u8c1_nearest, src, M, dsize, np.uint8, 1)


def render_weight_image(w, h):
    """The 8-bit weight image of Map2DRender::renderFrames (Map2DRender.cpp:507-529)."""
    out = np.zeros((h, w), np.uint8)
    lib().orc_render_weight_image(w, h, out.ctypes.data)
    return out


def weight_image_u8(w, h, weight_type=0):
    out = np.zeros((h, w), np.uint8)
    lib().orc_weight_image_u8(w, h, weight_type, out.ctypes.data)
    return out


def weight_image_f32(w, h, weight_type=0):
    out = np.zeros((h, w), np.float32)
    lib().orc_weight_image_f32(w, h, weight_type, out.ctypes.data)
    return out


# ---------------------------------------------------------------- the Map2D object
TYPE_CPU, TYPE_GPU, TYPE_MULTIBAND, TYPE_RENDER = 1, 2, 3, 4


class OracleMap2D:
    """Reference-shaped interface: Map2D::create / prepare / feed / save (Map2D.h:83-97)."""

    def __init__(self, type_, cfg=None, **kw):
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.type = TYPE_CPU if type_ == TYPE_GPU else type_
        self._h = C.c_void_p()
        rc = lib().orc_create(type_, C.byref(self.cfg), C.byref(self._h))
        if rc != 0:
            raise ValueError("orc_create failed: %d" % rc)
        self.levels = min(self.cfg.band_number if self.cfg.band_number > 0 else 5, 8) + 1

    @classmethod
    def create(cls, type_=TYPE_CPU, thread=False, **kw):
        return cls(type_, **kw)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def prepare(self, plane, camera, poses):
        plane = np.ascontiguousarray(plane, np.float64).reshape(7)
        camera = np.ascontiguousarray(camera, np.float64).reshape(6)
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        return lib().orc_prepare(self._h, _dptr(plane), _dptr(camera), len(poses), _dptr(poses)) == 0

    def feed(self, img, pose):
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.strides[2] == 1 and img.strides[1] == 3
        pose = np.ascontiguousarray(pose, np.float64).reshape(7)
        return lib().orc_feed(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], _dptr(pose)) == 0

    # ---- Map2DRender (type 4): one batch -> one blended canvas (Map2DRender.cpp:479-760)
    def render_frames(self, frames, poses):
        frames = np.ascontiguousarray(frames, np.uint8)
        assert frames.ndim == 4 and frames.shape[3] == 3
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n, h, w = frames.shape[:3]
        res = np.zeros(max(n, 1), np.int32)
        rc = lib().orc_render_frames(self._h, n, frames.ctypes.data, w * h * 3, w, h, w * 3, _dptr(poses),
                                     res.ctypes.data_as(C.POINTER(C.c_int)))
        return rc, res[:n]

    def render_get(self):
        """(result int16 HxWx3, mask u8 HxW, num_bands, (tile_x0, tile_y0)) of the last render_frames, or None."""
        w, h, nb, tx, ty = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        if lib().orc_render_get(self._h, None, None, C.byref(w), C.byref(h), C.byref(nb), C.byref(tx), C.byref(ty)) != 0:
            return None
        res = np.zeros((h.value, w.value, 3), np.int16)
        mask = np.zeros((h.value, w.value), np.uint8)
        lib().orc_render_get(self._h, res.ctypes.data, mask.ctypes.data, None, None, None, None, None)
        return res, mask, nb.value, (tx.value, ty.value)

    def render_warped(self, i):
        """(img u8 hxwx3, mask u8 hxw, (cx, cy)) handed to the blender for frame i, or None if the frame was skipped."""
        iw, ih, cx, cy = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        if lib().orc_render_warped(self._h, i, None, None, C.byref(iw), C.byref(ih), C.byref(cx), C.byref(cy)) != 0:
            return None
        img = np.zeros((ih.value, iw.value, 3), np.uint8)
        mask = np.zeros((ih.value, iw.value), np.uint8)
        lib().orc_render_warped(self._h, i, img.ctypes.data, mask.ctypes.data, C.byref(iw), C.byref(ih), C.byref(cx), C.byref(cy))
        return img, mask, (cx.value, cy.value)

    def grid(self):
        w, h, lp = C.c_int(), C.c_int(), C.c_double()
        mn, mx = np.zeros(3), np.zeros(3)
        rc = lib().orc_get_grid(self._h, C.byref(w), C.byref(h), _dptr(mn), _dptr(mx), C.byref(lp))
        if rc:
            raise RuntimeError("not prepared")
        return {"w": w.value, "h": h.value, "min": mn, "max": mx, "length_pixel": lp.value}

    def last_rect(self):
        r = (C.c_int * 4)()
        lib().orc_last_rect(self._h, r)
        return tuple(r)

    def get_tile(self, tx, ty, level=0):
        n = 256 >> level
        if self.type == TYPE_MULTIBAND:
            lap, wgt = np.zeros((n, n, 3), np.int16), np.zeros((n, n), np.float32)
            rc = lib().orc_get_tile(self._h, tx, ty, level, lap.ctypes.data, wgt.ctypes.data)
            return None if rc else (lap, wgt)
        out = np.zeros((256, 256, 4), np.uint8)
        rc = lib().orc_get_tile(self._h, tx, ty, 0, out.ctypes.data, None)
        return None if rc else out

    def get_tile_image(self, tx, ty, high_quality=True):
        out = np.zeros(256 * 256 * 4, np.uint8)
        cn = C.c_int()
        rc = lib().orc_get_tile_image(self._h, tx, ty, int(high_quality), out.ctypes.data, C.byref(cn))
        return None if rc else out[:256 * 256 * cn.value].reshape(256, 256, cn.value).copy()

    def get_image(self):
        w, h, cn, tx, ty = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = lib().orc_get_image(self._h, None, C.byref(w), C.byref(h), C.byref(cn), C.byref(tx), C.byref(ty))
        if rc:
            return None
        out = np.zeros((h.value, w.value, cn.value), np.uint8)
        lib().orc_get_image(self._h, out.ctypes.data, C.byref(w), C.byref(h), C.byref(cn), C.byref(tx), C.byref(ty))
        return out, (tx.value, ty.value)

    # --- CPU stand-in of the sharded library (world_size-2 gloo tests of the N>1 host logic) ------------------
    def tile_bytes(self):
        return int(lib().orc_tile_bytes(self._h))

    def tile_count(self):
        return int(lib().orc_tile_count(self._h))

    def export_tiles(self, dst_ptr, max_tiles, on_device=False):
        xy = np.zeros((max(max_tiles, 1), 2), np.int32)
        n = C.c_int()
        rc = lib().orc_export_tiles(self._h, max_tiles, xy.ctypes.data_as(C.POINTER(C.c_int)), dst_ptr, 0, C.byref(n))
        assert rc == 0
        return xy[:n.value].copy()

    def export_tiles_rect(self, rect_abs, dst_ptr, max_tiles, on_device=False):
        r = (C.c_int * 4)(*[int(v) for v in rect_abs])
        xy = np.zeros((max(max_tiles, 1), 2), np.int32)
        n = C.c_int()
        rc = lib().orc_export_tiles_rect(self._h, r, max_tiles, xy.ctypes.data_as(C.POINTER(C.c_int)), dst_ptr, 0, C.byref(n))
        assert rc == 0
        return n.value if max_tiles == 0 else xy[:n.value].copy()

    def drop_tiles_rect(self, rect_abs):
        r = (C.c_int * 4)(*[int(v) for v in rect_abs])
        n = C.c_int()
        lib().orc_drop_tiles_rect(self._h, r, C.byref(n))
        return n.value

    def tile_bbox(self):
        r = (C.c_int * 4)()
        return tuple(r) if lib().orc_tile_bbox(self._h, r) == 0 else None

    def get_image_rect(self, window_abs, crop_abs, out_ptr=None, on_device=False):
        win = (C.c_int * 4)(*[int(v) for v in window_abs])
        crop = (C.c_int * 4)(*[int(v) for v in crop_abs])
        w, h, cn = C.c_int(), C.c_int(), C.c_int()
        assert lib().orc_get_image_rect(self._h, None, 0, win, crop, C.byref(w), C.byref(h), C.byref(cn)) == 0
        if out_ptr is not None:
            assert lib().orc_get_image_rect(self._h, out_ptr, 0, win, crop, C.byref(w), C.byref(h), C.byref(cn)) == 0
            return h.value, w.value, cn.value
        out = np.empty((h.value, w.value, cn.value), np.uint8)
        assert lib().orc_get_image_rect(self._h, out.ctypes.data, 0, win, crop, C.byref(w), C.byref(h), C.byref(cn)) == 0
        return out

    def import_tiles(self, xy, src_ptr, on_device=False):
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        return lib().orc_import_tiles(self._h, len(xy), xy.ctypes.data_as(C.POINTER(C.c_int)), src_ptr, 0) == 0

    def plan_rects(self, poses):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        rects = np.zeros((len(poses), 4), np.int32)
        assert lib().orc_plan_rects(self._h, len(poses), _dptr(poses), rects.ctypes.data_as(C.POINTER(C.c_int))) == 0
        return rects

    def set_shard(self, rank, count, axis, span, origin=0):
        return lib().orc_set_shard(self._h, rank, count, axis, span, origin) == 0

    def feed_poses(self, poses):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        res = np.zeros(len(poses), np.int32)
        rc = lib().orc_feed_poses(self._h, len(poses), _dptr(poses), res.ctypes.data_as(C.POINTER(C.c_int)))
        if rc < 0:
            raise RuntimeError("feed_poses: a pose touches tiles this shard owns")
        return res

    def feed_batch(self, base_ptr, n, frame_stride, w, h, stride, poses, on_device=False):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        res = np.zeros(n, np.int32)
        for k in range(n):
            res[k] = lib().orc_feed(self._h, base_ptr + k * frame_stride, w, h, stride, _dptr(poses[k]))
        return res

    def sync(self):
        return True

    def stats(self):
        s = Stats()
        lib().orc_get_stats(self._h, C.byref(s))
        return s.as_dict()

    def compute_bounds(self, poses):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        rects = np.zeros((n, 4), np.int32)
        hinv = np.zeros((n, 9), np.float64)
        lib().orc_compute_bounds(self._h, n, _dptr(poses), rects.ctypes.data_as(C.POINTER(C.c_int)), _dptr(hinv))
        return rects, hinv.reshape(n, 3, 3)
