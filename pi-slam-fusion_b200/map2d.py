"""Host-side mirror of the reference's Map2D plugin interface (Map2DFusion/Map2D.h:79-98) on top of the C-ABI
library libmap2d_b200.so (include/map2d_b200.h).

    m = Map2D.create(Map2D.TypeMultiBandCPU, thread=False)     # Map2D::create      Map2D.cpp:51-66
    m.prepare(plane, camera, poses)                           # Map2D::prepare     Map2D.h:88
    m.feed(img_bgr_u8, pose_c2w)                              # Map2D::feed        Map2D.h:91
    m.save("result.png")                                      # Map2D::save        Map2D.h:95
    m.queueSize()                                             # Map2D::queueSize   Map2D.h:97

Same names, argument meaning and return convention (bool) as the reference.  There is no CPU path: if the shared
library or a CUDA device is missing, importing works but create() raises.  This module never imports oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmap2d_b200.so")
MAX_LEVELS = 9
OK, REJECTED = 0, 1
ERR_ARG, ERR_STATE, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_IO = -1, -2, -3, -4, -5, -6


class Config(C.Structure):
    """m2d_config"""
    _fields_ = [("scale", C.c_double), ("resolution", C.c_double), ("weight_type", C.c_int),
                ("band_number", C.c_int), ("force_float", C.c_int), ("background", C.c_int),
                ("thread", C.c_int), ("device", C.c_int), ("shard_rank", C.c_int), ("shard_count", C.c_int),
                ("shard_axis", C.c_int), ("shard_span", C.c_int), ("collect_stats", C.c_int),
                ("batch_frames", C.c_int), ("f32_mode", C.c_int), ("render_blend", C.c_int),
                ("render_bands", C.c_int)]


class Stats(C.Structure):
    """m2d_stats"""
    _fields_ = [("frames_fed", C.c_uint64), ("frames_fused", C.c_uint64), ("input_px", C.c_uint64),
                ("region_px", C.c_uint64 * MAX_LEVELS), ("fresh_px", C.c_uint64 * MAX_LEVELS),
                ("win_px", C.c_uint64 * MAX_LEVELS), ("footprint_px", C.c_uint64), ("need_px", C.c_uint64 * MAX_LEVELS),
                ("needw_px", C.c_uint64 * MAX_LEVELS)]

    def as_dict(self):
        return {"frames_fed": self.frames_fed, "frames_fused": self.frames_fused, "input_px": self.input_px,
                "region_px": list(self.region_px), "fresh_px": list(self.fresh_px), "win_px": list(self.win_px),
                "footprint_px": self.footprint_px, "need_px": list(self.need_px), "needw_px": list(self.needw_px)}


EXPORTS = ["m2d_config_default", "m2d_create", "m2d_create_multi", "m2d_destroy", "m2d_prepare", "m2d_feed", "m2d_feed_device",
           "m2d_feed_batch", "m2d_feed_batch_ptrs", "m2d_device_alloc", "m2d_device_free", "m2d_ipc_export", "m2d_ipc_open", "m2d_ipc_close", "m2d_feed_poses", "m2d_plan_rects", "m2d_set_shard", "m2d_sync", "m2d_queue_size", "m2d_set_stream", "m2d_reset", "m2d_set_input_event", "m2d_get_grid",
           "m2d_last_rect", "m2d_get_tile", "m2d_get_image", "m2d_save", "m2d_tile_bytes", "m2d_tile_state_bytes", "m2d_tile_count",
           "m2d_export_tiles", "m2d_import_tiles", "m2d_export_tiles_rect", "m2d_drop_tiles_rect", "m2d_tile_bbox", "m2d_get_image_rect", "m2d_poll_changed", "m2d_get_tile_image", "m2d_save_state", "m2d_load_state", "m2d_get_stats", "m2d_last_error",
           "m2d_launch_count", "m2d_profile", "m2d_get_kernel_times", "m2d_alloc_host", "m2d_free_host", "m2d_compute_bounds",
           "m2d_tile_gps_corners", "m2d_reach_table", "m2d_weight_reach_table", "m2d_cell_weight_bounds", "m2d_ingest_open", "m2d_ingest_open_seeded", "m2d_ingest_abort", "m2d_ingest_push", "m2d_ingest_pause", "m2d_ingest_drain", "m2d_ingest_close", "m2d_ingest_stats",
           "m2d_render_frames", "m2d_render_get", "m2d_pull_cell_rect"]

_lib = None


def lib():
    """Load libmap2d_b200.so (built by __graft_entry__.build() / csrc/Makefile).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmap2d_b200.so is not built (%s); run `python __graft_entry__.py` or "
                           "`make -C pi-slam-fusion_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.m2d_config_default.argtypes = [C.POINTER(Config)]
    L.m2d_config_default.restype = None
    L.m2d_create.argtypes = [C.c_int, C.POINTER(Config), C.POINTER(vp)]
    L.m2d_create_multi.argtypes = [C.c_int, C.POINTER(Config), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
    L.m2d_destroy.argtypes = [vp]
    L.m2d_destroy.restype = None
    L.m2d_prepare.argtypes = [vp, dp, dp, C.c_int, dp]
    L.m2d_feed.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, dp]
    L.m2d_feed_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, dp]
    L.m2d_feed_batch.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_size_t, dp, C.c_int, ip]
    L.m2d_feed_batch_ptrs.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_size_t, dp, C.c_int, ip]
    L.m2d_device_alloc.argtypes = [C.c_int, C.c_size_t]
    L.m2d_device_alloc.restype = vp
    L.m2d_device_free.argtypes = [C.c_int, vp]
    L.m2d_device_free.restype = None
    L.m2d_ipc_export.argtypes = [vp, C.c_char_p]
    L.m2d_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.m2d_ipc_close.argtypes = [C.c_int, vp]
    L.m2d_feed_poses.argtypes = [vp, C.c_int, dp, ip]
    L.m2d_plan_rects.argtypes = [vp, C.c_int, dp, ip]
    L.m2d_tile_gps_corners.argtypes = [dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, dp]
    L.m2d_reach_table.argtypes = [C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte)]
    L.m2d_weight_reach_table.argtypes = [C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_ubyte)]
    L.m2d_cell_weight_bounds.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.m2d_ingest_open.argtypes = [vp, C.c_int, C.c_int]
    L.m2d_ingest_open_seeded.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.m2d_ingest_abort.argtypes = [vp]
    L.m2d_ingest_push.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_int, dp]
    L.m2d_ingest_pause.argtypes = [vp, C.c_int]
    L.m2d_ingest_drain.argtypes = [vp]
    L.m2d_ingest_close.argtypes = [vp]
    u64p = C.POINTER(C.c_uint64)
    L.m2d_ingest_stats.argtypes = [vp, u64p, u64p, u64p, u64p]
    L.m2d_set_shard.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.m2d_sync.argtypes = [vp]
    L.m2d_queue_size.argtypes = [vp]
    L.m2d_set_stream.argtypes = [vp, vp]
    L.m2d_reset.argtypes = [vp]
    L.m2d_set_input_event.argtypes = [vp, vp]
    L.m2d_get_grid.argtypes = [vp, ip, ip, dp, dp, dp]
    L.m2d_last_rect.argtypes = [vp, ip]
    L.m2d_get_tile.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.m2d_get_image.argtypes = [vp, vp, ip, ip, ip, ip, ip]
    L.m2d_save.argtypes = [vp, C.c_char_p]
    L.m2d_tile_bytes.argtypes = [vp]
    L.m2d_tile_bytes.restype = C.c_size_t
    L.m2d_tile_state_bytes.argtypes = [vp]
    L.m2d_tile_state_bytes.restype = C.c_size_t
    L.m2d_tile_count.argtypes = [vp]
    L.m2d_export_tiles.argtypes = [vp, C.c_int, ip, vp, C.c_int, ip]
    L.m2d_import_tiles.argtypes = [vp, C.c_int, ip, vp, C.c_int]
    L.m2d_export_tiles_rect.argtypes = [vp, ip, C.c_int, ip, vp, C.c_int, ip]
    L.m2d_drop_tiles_rect.argtypes = [vp, ip, ip]
    L.m2d_tile_bbox.argtypes = [vp, ip]
    L.m2d_get_image_rect.argtypes = [vp, vp, C.c_int, ip, ip, ip, ip, ip]
    L.m2d_poll_changed.argtypes = [vp, C.c_int, ip, ip]
    L.m2d_get_tile_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, ip]
    L.m2d_save_state.argtypes = [vp, C.c_char_p]
    L.m2d_load_state.argtypes = [vp, C.c_char_p]
    L.m2d_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.m2d_last_error.argtypes = [vp]
    L.m2d_last_error.restype = C.c_char_p
    L.m2d_launch_count.argtypes = [vp]
    L.m2d_launch_count.restype = C.c_uint64
    L.m2d_profile.argtypes = [vp, C.c_int]
    L.m2d_get_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_uint64)]
    L.m2d_alloc_host.argtypes = [C.c_size_t]
    L.m2d_alloc_host.restype = vp
    L.m2d_free_host.argtypes = [vp]
    L.m2d_free_host.restype = None
    L.m2d_compute_bounds.argtypes = [vp, C.c_int, dp, ip, dp]
    L.m2d_render_frames.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_size_t, dp, C.c_int, ip]
    L.m2d_render_get.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip]
    L.m2d_pull_cell_rect.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, ip]
    _lib = L
    return L


def default_config(**kw):
    c = Config()
    lib().m2d_config_default(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def tile_gps_corners(plane, grid, tx, ty, gps_origin):
    """m2d_tile_gps_corners for a grid dict as returned by Map2D.grid() (works without a GPU: pure host arithmetic)."""
    plane = np.ascontiguousarray(plane, np.float64).reshape(7)
    org = np.ascontiguousarray(np.asarray(gps_origin, np.float64).reshape(-1)[:2])
    tl, br = np.zeros(3), np.zeros(3)
    rc = lib().m2d_tile_gps_corners(_dptr(plane), float(grid["min"][0]), float(grid["min"][1]), 256.0 * float(grid["length_pixel"]),
                                    int(tx), int(ty), _dptr(org), _dptr(tl), _dptr(br))
    if rc != OK:
        raise RuntimeError("m2d_tile_gps_corners failed: %d" % rc)
    return tl, br


def reach_table(levels):
    """(lo, hi) 6x6 uint8 arrays of m2d_reach_table: cells of Gaussian level k a level-m winner depends on."""
    lo, hi = np.zeros(36, np.uint8), np.zeros(36, np.uint8)
    rc = lib().m2d_reach_table(levels, lo.ctypes.data_as(C.POINTER(C.c_ubyte)), hi.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if rc != OK:
        raise ValueError("m2d_reach_table(%d) failed: %d" % (levels, rc))
    return lo.reshape(6, 6), hi.reshape(6, 6)


def weight_reach_table(levels):
    """(lo, hi) 6x6 uint8 arrays of m2d_weight_reach_table: cells of weight level k a competitive cell of level m depends on."""
    lo, hi = np.zeros(36, np.uint8), np.zeros(36, np.uint8)
    rc = lib().m2d_weight_reach_table(levels, lo.ctypes.data_as(C.POINTER(C.c_ubyte)), hi.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if rc != OK:
        raise ValueError("m2d_weight_reach_table(%d) failed: %d" % (levels, rc))
    return lo.reshape(6, 6), hi.reshape(6, 6)


def cell_weight_bounds(hinv, nx, ny, sw, sh, weight_type, level, cx, cy):
    """(lo, hi) of m2d_cell_weight_bounds: closed-form bounds of a frame's level-`level` weight over one 32-px cell."""
    hinv = np.ascontiguousarray(hinv, np.float64).reshape(9)
    lo, hi = C.c_float(), C.c_float()
    rc = lib().m2d_cell_weight_bounds(_dptr(hinv), nx, ny, sw, sh, weight_type, level, cx, cy, C.byref(lo), C.byref(hi))
    if rc != OK:
        raise ValueError("m2d_cell_weight_bounds failed: %d" % rc)
    return lo.value, hi.value


def pull_cell_rect(hinv, X0, Y0, sw, sh):
    """(ok, (lox, hix, loy, hiy)) of m2d_pull_cell_rect: the source px pull mode fetches for the cell at region px (X0, Y0)."""
    hinv = np.ascontiguousarray(hinv, np.float64).reshape(9)
    r = (C.c_int * 4)()
    rc = lib().m2d_pull_cell_rect(_dptr(hinv), int(X0), int(Y0), int(sw), int(sh), r)
    if rc < 0:
        raise ValueError("m2d_pull_cell_rect failed: %d" % rc)
    return rc == OK, tuple(r)


def map2d_update_command(plane, grid, tx, ty, gps_origin, image_name="LastTexMat"):
    """`Map2DUpdate LastTexMat <tl> <br>` exactly as MultiBandMap2DCPU.cpp:754-755 formats it (fixed, 9 decimals,
    Point3d streamed as `x y z`)."""
    tl, br = tile_gps_corners(plane, grid, tx, ty, gps_origin)
    return "Map2DUpdate %s %s %s" % (image_name, " ".join("%.9f" % v for v in tl), " ".join("%.9f" % v for v in br))


def tile_overlay(map2d, tx, ty, high_quality=True):
    """The pixmap Map2DItemHandle builds from (LastTexMat, LastTexMatWeight): rows upside-down; multi-band -> B,G,R,A
    bytes (QImage::Format_ARGB32) with A = 255 where the level-0 weight is non-zero (Map2DItem.cpp:68-84); weighted
    (no float weight image) -> R,G,B bytes (Map2DItem.cpp:56-67).  `map2d`: anything with get_tile_image/get_tile."""
    img = map2d.get_tile_image(tx, ty, high_quality)
    if img is None:
        return None
    if img.shape[2] == 3:
        w0 = map2d.get_tile(tx, ty, 0)[1]
        out = np.dstack([img, np.where(w0 != 0, 255, 0).astype(np.uint8)])
        return np.ascontiguousarray(out[::-1])
    return np.ascontiguousarray(img[::-1, :, 2::-1])


class DeviceBuffer:
    """A cudaMalloc'ed buffer that other processes on the node can map (m2d_device_alloc / m2d_ipc_*).  `.tensor()` wraps it
    as a torch uint8 tensor (no copy, via __cuda_array_interface__) so that it can be filled like any other tensor."""

    def __init__(self, device, nbytes):
        self.device, self.nbytes = int(device), int(nbytes)
        self.ptr = lib().m2d_device_alloc(self.device, max(self.nbytes, 16))
        if not self.ptr:
            raise MemoryError("m2d_device_alloc(%d bytes) failed on device %d" % (nbytes, device))
        self.__cuda_array_interface__ = {"shape": (max(self.nbytes, 16),), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device=torch.device("cuda", self.device))[:self.nbytes]

    def export(self):
        h = C.create_string_buffer(64)
        if lib().m2d_ipc_export(self.ptr, h) != OK:
            raise RuntimeError("m2d_ipc_export failed")
        return h.raw

    def free(self):
        if self.ptr:
            lib().m2d_device_free(self.device, self.ptr)
            self.ptr = None


def ipc_open(device, handle):
    """Map a buffer another process exported with DeviceBuffer.export(); returns its address in this process."""
    p = C.c_void_p()
    if lib().m2d_ipc_open(int(device), handle, C.byref(p)) != OK:
        raise RuntimeError("m2d_ipc_open failed (CUDA IPC / peer access not available between these GPUs?)")
    return p.value


def ipc_close(device, ptr):
    lib().m2d_ipc_close(int(device), ptr)


def pinned_empty(shape, dtype=np.uint8):
    """numpy array backed by page-locked host memory from m2d_alloc_host (truly asynchronous feeds)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().m2d_alloc_host(n)
    if not p:
        raise MemoryError("m2d_alloc_host(%d) failed" % n)
    buf = (C.c_uint8 * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p  # free with free_pinned(p) once arr is no longer used


def free_pinned(p):
    lib().m2d_free_host(p)


class Map2D:
    """Reference-shaped object; see module docstring."""
    NoType, TypeCPU, TypeGPU, TypeMultiBandCPU, TypeRender = 0, 1, 2, 3, 4  # Map2D.h:83

    def __init__(self, type_, cfg, devices=None):
        self.cfg = cfg
        self.type = self.TypeCPU if type_ == self.TypeGPU else type_
        self._h = C.c_void_p()
        if devices is not None and len(devices) > 1:   # one process, several GPUs (m2d_create_multi)
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = lib().m2d_create_multi(type_, C.byref(cfg), len(devices), arr, C.byref(self._h))
        else:
            if devices:
                cfg.device = int(devices[0])
            rc = lib().m2d_create(type_, C.byref(cfg), C.byref(self._h))
        if rc != OK:
            raise RuntimeError("m2d_create(type=%d) failed with status %d (no CUDA device? unsupported type?)" % (type_, rc))
        self.levels = (min(cfg.band_number if cfg.band_number > 0 else 5, 8) + 1) if self.type == self.TypeMultiBandCPU else 1

    @classmethod
    def create(cls, type_=1, thread=True, **kw):
        """Map2D::create(type, thread) — returns None for NoType like the reference's null SPtr."""
        if type_ == cls.NoType:
            return None
        devices = kw.pop("devices", None)
        cfg = kw.pop("cfg", None) or default_config(thread=int(bool(thread)), **kw)
        return cls(type_, cfg, devices)

    def close(self):
        if getattr(self, "_h", None):
            lib().m2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError("map2d_b200 error %d: %s" % (rc, lib().m2d_last_error(self._h).decode()))
        return rc == OK

    # --- Map2D interface -----------------------------------------------------------------------------
    def prepare(self, plane, camera, frames):
        """frames: sequence of poses (n x 7) or of (img, pose) pairs like the reference's deque.
        thread=True (Map2D::create's second argument) behaves like the reference: prepare() starts the worker, whose
        queue is seeded with the prepare-frames themselves (Map2D.cpp:42, Map2DCPU.cpp:384-413) -- they are rendered
        first, in order -- and feed() only enqueues (drop-oldest beyond 20, Map2DCPU.cpp:139-142); a second prepare()
        discards what was still queued for the old map.  thread=False never renders the prepare-frames."""
        pairs = [f for f in frames if isinstance(f, (tuple, list)) and len(f) == 2 and np.ndim(f[1]) == 1]
        poses = [f[1] if isinstance(f, (tuple, list)) and len(f) == 2 and np.ndim(f[1]) == 1 else f for f in frames]
        poses = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 7))
        plane = np.ascontiguousarray(plane, np.float64).reshape(7)
        camera = np.ascontiguousarray(camera, np.float64).reshape(6)
        threaded = bool(self.cfg.thread)
        if threaded:
            lib().m2d_ingest_abort(self._h)
        if not self._check(lib().m2d_prepare(self._h, _dptr(plane), _dptr(camera), len(poses), _dptr(poses))):
            return False
        if self.type == self.TypeRender:
            # Map2DRender: the worker started by prepare() takes everything queued -- the prepare-frames -- as ONE batch through
            # renderFrames and stops (Map2DRender.cpp:419-420, 469-478, 758, 812-829); thread=False never renders (:464-467)
            imgs = [np.asarray(f[0]) for f in pairs]
            if threaded and imgs:
                self.render_frames(np.stack(imgs), np.asarray([f[1] for f in pairs], np.float64))
            return True
        if not threaded:
            return True
        self._check(lib().m2d_ingest_open_seeded(self._h, 20, len(pairs), 1))
        for img, pose in pairs:
            img = np.asarray(img)
            if img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3:
                self.ingest_push(np.ascontiguousarray(img), pose)
        return self.ingest_pause(False)

    def feed(self, img, pose):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3 or img.strides[2] != 1 or img.strides[1] != 3:
            print("Map2DB200::feed: image must be CV_8UC3")  # reference: type()!=CV_8UC3 -> false
            return False
        pose = np.ascontiguousarray(pose, np.float64).reshape(7)
        if self.type == self.TypeRender and self.cfg.thread:
            return True   # Map2DRender::feed only queues (Map2DRender.cpp:446-452); its worker has stopped after the first batch
        if self.cfg.thread and self._ingest_open():
            return self._check(lib().m2d_ingest_push(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], 3, _dptr(pose)))
        return self._check(lib().m2d_feed(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], _dptr(pose)))

    def _ingest_open(self):
        v = C.c_uint64()
        return lib().m2d_ingest_stats(self._h, C.byref(v), None, None, None) == OK

    def feed_device(self, dev_ptr, w, h, stride, pose):
        pose = np.ascontiguousarray(pose, np.float64).reshape(7)
        return self._check(lib().m2d_feed_device(self._h, dev_ptr, w, h, stride, _dptr(pose)))

    def feed_batch(self, base_ptr, n, frame_stride, w, h, stride, poses, on_device):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        res = np.zeros(n, np.int32)
        rc = lib().m2d_feed_batch(self._h, n, base_ptr, frame_stride, w, h, stride, _dptr(poses), int(on_device),
                                  res.ctypes.data_as(C.POINTER(C.c_int)))
        self._check(rc)
        return res

    def feed_batch_ptrs(self, ptrs, w, h, stride, poses, on_device=True):
        """feed_batch with one address per frame (e.g. some frames mapped from a neighbouring GPU's memory)."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(ptrs)
        arr = (C.c_void_p * max(n, 1))(*[int(p) for p in ptrs])
        res = np.zeros(n, np.int32)
        self._check(lib().m2d_feed_batch_ptrs(self._h, n, arr, w, h, stride, _dptr(poses), int(on_device), res.ctypes.data_as(C.POINTER(C.c_int))))
        return res

    # --- Map2DRender (type 4): one batch -> one blended canvas (Map2DRender.cpp:479-760) ---------------------
    def render_frames(self, frames, poses, on_device=False, w=None, h=None):
        """frames: n x H x W x 3 uint8 array (host), or a device address of n packed frames with on_device=True (+ w, h).
        Returns the per-frame status array (0 = blended, 1 = skipped like the reference: oblique view)."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        if on_device:
            ptr = int(frames)
        else:
            frames = np.ascontiguousarray(frames, np.uint8)
            assert frames.ndim == 4 and frames.shape[0] == n and frames.shape[3] == 3
            h, w = frames.shape[1:3]
            ptr = frames.ctypes.data
        res = np.zeros(max(n, 1), np.int32)
        rc = lib().m2d_render_frames(self._h, n, ptr, w * h * 3, w, h, w * 3, _dptr(poses), int(on_device), res.ctypes.data_as(C.POINTER(C.c_int)))
        self._check(rc)
        return rc, res[:n]

    def render_get(self):
        """(result int16 HxWx3, mask uint8 HxW, num_bands, (tile_x0, tile_y0)) of the last render_frames, or None."""
        w, h, nb, tx, ty = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        if lib().m2d_render_get(self._h, None, None, C.byref(w), C.byref(h), C.byref(nb), C.byref(tx), C.byref(ty)) != OK:
            return None
        res = np.zeros((h.value, w.value, 3), np.int16)
        mask = np.zeros((h.value, w.value), np.uint8)
        self._check(lib().m2d_render_get(self._h, res.ctypes.data, mask.ctypes.data, None, None, None, None, None))
        return res, mask, nb.value, (tx.value, ty.value)

    def feed_poses(self, poses):
        """Sharded runs: frames whose pixels this shard does not need (grid growth only), see m2d_feed_poses."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        res = np.zeros(len(poses), np.int32)
        self._check(lib().m2d_feed_poses(self._h, len(poses), _dptr(poses), res.ctypes.data_as(C.POINTER(C.c_int))))
        return res

    def plan_rects(self, poses):
        """Absolute tile rect of every frame as sequential feeds will compute it (dry run), see m2d_plan_rects."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        rects = np.zeros((len(poses), 4), np.int32)
        self._check(lib().m2d_plan_rects(self._h, len(poses), _dptr(poses), rects.ctypes.data_as(C.POINTER(C.c_int))))
        return rects

    # --- ingest seam: bounded drop-oldest queue + worker in front of feed() --------------
    def ingest_open(self, capacity=30, start_paused=False):
        return self._check(lib().m2d_ingest_open(self._h, capacity, int(start_paused)))

    def ingest_push(self, img, pose):
        """img: HxWx3 (BGR) or HxWx4 (BGRA) uint8.  Never blocks; a full queue drops its oldest frame."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4) or img.strides[2] != 1 or img.strides[1] != img.shape[2]:
            raise ValueError("ingest_push needs a uint8 HxWx3 or HxWx4 image with packed pixels")
        pose = np.ascontiguousarray(pose, np.float64).reshape(7)
        return self._check(lib().m2d_ingest_push(self._h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], img.shape[2], _dptr(pose)))

    def ingest_pause(self, paused=True):
        return self._check(lib().m2d_ingest_pause(self._h, int(paused)))

    def ingest_drain(self):
        return self._check(lib().m2d_ingest_drain(self._h))

    def ingest_close(self):
        return self._check(lib().m2d_ingest_close(self._h))

    def ingest_abort(self):
        """Close without draining: queued frames are discarded (what a second prepare() does to the old queue)."""
        return self._check(lib().m2d_ingest_abort(self._h))

    def ingest_open_seeded(self, capacity=20, seed_frames=0, start_paused=True):
        return self._check(lib().m2d_ingest_open_seeded(self._h, capacity, seed_frames, int(start_paused)))

    def ingest_stats(self):
        v = [C.c_uint64() for _ in range(4)]
        self._check(lib().m2d_ingest_stats(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("pushed", "dropped", "fed", "fused"), (int(x.value) for x in v)))

    def set_shard(self, rank, count, axis, span, origin=0):
        """Re-partition tile ownership (only while the map holds no tiles), see m2d_set_shard."""
        return self._check(lib().m2d_set_shard(self._h, rank, count, axis, span, origin))

    def save(self, filename):
        if self.cfg.thread and self._ingest_open():
            self.ingest_drain()
        return self._check(lib().m2d_save(self._h, os.fsencode(filename)))

    def queueSize(self):
        return lib().m2d_queue_size(self._h)

    def draw(self):  # GL display is out of scope; kept so the interface is complete
        return None

    # --- additions over the reference (in-memory getters) -----------------------------
    def sync(self):
        return self._check(lib().m2d_sync(self._h))

    def reset(self):
        return self._check(lib().m2d_reset(self._h))

    def set_input_event(self, cuda_event):
        """The next feed call's pixel-reading kernels wait for this cudaEvent_t (e.g. torch.cuda.Event().cuda_event)."""
        return self._check(lib().m2d_set_input_event(self._h, cuda_event))

    def set_stream(self, cuda_stream):
        return self._check(lib().m2d_set_stream(self._h, cuda_stream))

    def grid(self):
        w, h, lp = C.c_int(), C.c_int(), C.c_double()
        mn, mx = np.zeros(3), np.zeros(3)
        rc = lib().m2d_get_grid(self._h, C.byref(w), C.byref(h), _dptr(mn), _dptr(mx), C.byref(lp))
        if rc:
            raise RuntimeError("not prepared")
        return {"w": w.value, "h": h.value, "min": mn, "max": mx, "length_pixel": lp.value}

    def last_rect(self):
        r = (C.c_int * 4)()
        lib().m2d_last_rect(self._h, r)
        return tuple(r)

    def get_tile(self, tx, ty, level=0):
        n = 256 >> level
        if self.type == self.TypeMultiBandCPU:
            lap, wgt = np.zeros((n, n, 3), np.int16), np.zeros((n, n), np.float32)
            rc = lib().m2d_get_tile(self._h, tx, ty, level, lap.ctypes.data, wgt.ctypes.data)
            return (lap, wgt) if self._check(rc) else None
        out = np.zeros((256, 256, 4), np.uint8)
        rc = lib().m2d_get_tile(self._h, tx, ty, 0, out.ctypes.data, None)
        return out if self._check(rc) else None

    def get_image(self, out=None):
        """In-memory save(): (image, (tile_min_x, tile_min_y)).  `out`: optional preallocated uint8 buffer (e.g. from
        pinned_empty) of at least h*w*channels bytes; a view of it is returned."""
        w, h, cn, tx, ty = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        for _ in range(8):   # more than one round only while an ingest worker is growing the map under us
            rc = lib().m2d_get_image(self._h, None, C.byref(w), C.byref(h), C.byref(cn), C.byref(tx), C.byref(ty))
            if not self._check(rc):
                return None
            nbytes = h.value * w.value * cn.value
            buf = np.empty(nbytes, np.uint8) if out is None else out
            flat = buf.reshape(-1)
            assert flat.dtype == np.uint8 and flat.size >= nbytes and flat.flags["C_CONTIGUOUS"]
            rc = lib().m2d_get_image(self._h, flat.ctypes.data, C.byref(w), C.byref(h), C.byref(cn), C.byref(tx), C.byref(ty))
            if rc == ERR_STATE:
                continue
            self._check(rc)
            return flat[:nbytes].reshape(h.value, w.value, cn.value), (tx.value, ty.value)
        raise RuntimeError("get_image: the mosaic kept growing; pause or drain the ingest queue first")

    # --- sharded runs: final tile gather -----------------------------------------------------------------
    def tile_bytes(self):
        return int(lib().m2d_tile_bytes(self._h))

    def tile_state_bytes(self):
        """Leading bytes of a tile record that are reference state (the rest is library-private), see m2d_tile_state_bytes."""
        return int(lib().m2d_tile_state_bytes(self._h))

    def tile_count(self):
        return int(lib().m2d_tile_count(self._h))

    def export_tiles(self, dst_ptr, max_tiles, on_device):
        """Copy every tile this shard holds to dst_ptr; returns their absolute tile coordinates (n x 2 int32)."""
        xy = np.zeros((max(max_tiles, 1), 2), np.int32)
        n = C.c_int()
        self._check(lib().m2d_export_tiles(self._h, max_tiles, xy.ctypes.data_as(C.POINTER(C.c_int)), dst_ptr, int(on_device), C.byref(n)))
        return xy[:n.value].copy()

    def export_tiles_rect(self, rect_abs, dst_ptr, max_tiles, on_device):
        """export_tiles for the tiles inside rect_abs = (x0, y0, x1, y1), absolute tile coordinates; max_tiles = 0 counts only
        (returns the count)."""
        r = (C.c_int * 4)(*[int(v) for v in rect_abs])
        xy = np.zeros((max(max_tiles, 1), 2), np.int32)
        n = C.c_int()
        self._check(lib().m2d_export_tiles_rect(self._h, r, max_tiles, xy.ctypes.data_as(C.POINTER(C.c_int)), dst_ptr, int(on_device), C.byref(n)))
        return n.value if max_tiles == 0 else xy[:n.value].copy()

    def drop_tiles_rect(self, rect_abs):
        r = (C.c_int * 4)(*[int(v) for v in rect_abs])
        n = C.c_int()
        self._check(lib().m2d_drop_tiles_rect(self._h, r, C.byref(n)))
        return n.value

    def tile_bbox(self):
        """(x0, y0, x1, y1) of the tiles this handle holds, absolute tile coordinates; None if it holds none."""
        r = (C.c_int * 4)()
        rc = lib().m2d_tile_bbox(self._h, r)
        return tuple(r) if self._check(rc) else None

    def get_image_rect(self, window_abs, crop_abs, out_ptr=None, on_device=False):
        """The collapse of save() over an explicit window of tiles, cropped (m2d_get_image_rect).  Without out_ptr: a
        new numpy array [h, w, channels]; with out_ptr (host or device memory): (h, w, channels)."""
        win = (C.c_int * 4)(*[int(v) for v in window_abs])
        crop = (C.c_int * 4)(*[int(v) for v in crop_abs])
        w, h, cn = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().m2d_get_image_rect(self._h, None, 0, win, crop, C.byref(w), C.byref(h), C.byref(cn)))
        if out_ptr is not None:
            self._check(lib().m2d_get_image_rect(self._h, out_ptr, int(on_device), win, crop, C.byref(w), C.byref(h), C.byref(cn)))
            return h.value, w.value, cn.value
        out = np.empty((h.value, w.value, cn.value), np.uint8)
        self._check(lib().m2d_get_image_rect(self._h, out.ctypes.data, 0, win, crop, C.byref(w), C.byref(h), C.byref(cn)))
        return out

    def import_tiles(self, xy, src_ptr, on_device):
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        return self._check(lib().m2d_import_tiles(self._h, len(xy), xy.ctypes.data_as(C.POINTER(C.c_int)), src_ptr, int(on_device)))

    # --- display path without GL: changed tiles and their textures (Map2D::draw) ------------------------------
    def poll_changed(self, max_tiles=65536):
        """Tiles touched since the last poll (the reference's Ele::Ischanged), as (tx, ty) grid coordinates."""
        xy = np.zeros((max_tiles, 2), np.int32)
        n = C.c_int()
        self._check(lib().m2d_poll_changed(self._h, max_tiles, xy.ctypes.data_as(C.POINTER(C.c_int)), C.byref(n)))
        return [tuple(int(v) for v in r) for r in xy[:n.value]]

    def get_tile_image(self, tx, ty, high_quality=True):
        """One tile as the reference would texture it (Ele::blend with neighbour borders when high_quality)."""
        out = np.zeros(256 * 256 * 4, np.uint8)
        cn = C.c_int()
        rc = lib().m2d_get_tile_image(self._h, tx, ty, int(high_quality), out.ctypes.data, C.byref(cn))
        if not self._check(rc):
            return None
        return out[:256 * 256 * cn.value].reshape(256, 256, cn.value).copy()

    # --- Map2DUpdate: the Google-map overlay of the display loop (MultiBandMap2DCPU.cpp:744-757, Map2DItem.cpp:36-99) ---
    def tile_gps_corners(self, tx, ty, plane, gps_origin):
        """({lng,lat,0} of the tile's top-left, of its bottom-right) for GPS.Origin = (lng, lat[, alt])."""
        return tile_gps_corners(plane, self.grid(), tx, ty, gps_origin)

    def map2d_update_command(self, tx, ty, plane, gps_origin, image_name="LastTexMat"):
        """The command string the reference sends to the MapWidget for an updated interior tile."""
        return map2d_update_command(plane, self.grid(), tx, ty, gps_origin, image_name)

    def tile_overlay(self, tx, ty, high_quality=True):
        return tile_overlay(self, tx, ty, high_quality)

    # --- checkpoint / resume ----------------------------------------------------------------------------
    def save_state(self, filename):
        return self._check(lib().m2d_save_state(self._h, os.fsencode(filename)))

    def load_state(self, filename):
        return self._check(lib().m2d_load_state(self._h, os.fsencode(filename)))

    def stats(self):
        s = Stats()
        self._check(lib().m2d_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def launch_count(self):
        return int(lib().m2d_launch_count(self._h))

    KERNEL_CLASSES = ("weighted_fuse", "mb_warp", "mb_pyrdown", "mb_select", "collapse", "misc", "mb_pyrtail",
                      "mbw_warp", "mbw_pyramid", "mbs_decide", "mbs_propagate", "mbs_warp", "mbs_pyramid", "mbs_lap", "mbc_bounds", "render")

    def profile(self, enable):
        return self._check(lib().m2d_profile(self._h, int(enable)))

    def kernel_times(self):
        """{class: (total_ms, launches)} accumulated since the last call (synchronises)."""
        ms = np.zeros(len(self.KERNEL_CLASSES), np.float64)
        cnt = np.zeros(len(self.KERNEL_CLASSES), np.uint64)
        self._check(lib().m2d_get_kernel_times(self._h, _dptr(ms), cnt.ctypes.data_as(C.POINTER(C.c_uint64))))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_CLASSES) if cnt[i]}

    def compute_bounds(self, poses):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        rects = np.zeros((n, 4), np.int32)
        hinv = np.zeros((n, 9), np.float64)
        self._check(lib().m2d_compute_bounds(self._h, n, _dptr(poses), rects.ctypes.data_as(C.POINTER(C.c_int)), _dptr(hinv)))
        return rects, hinv.reshape(n, 3, 3)
