"""The C-ABI library loads on a box without a GPU, exports every symbol include/map2d_b200.h declares, agrees with
the Python mirror on struct layout, and FAILS LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import pi_slam_fusion_b200.map2d as m2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "map2d_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m2d_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(m2d.LIB_PATH), "run `python __graft_entry__.py` (build()) first"
    m2d.lib()


def test_every_declared_symbol_is_exported():
    L = m2d.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "symbol %s declared in map2d_b200.h is not exported" % s
    assert sorted(m2d.EXPORTS) == syms, "python mirror out of sync with the header"


def test_only_m2d_symbols_have_default_visibility_prefix():
    out = subprocess.run(["nm", "-D", "--defined-only", m2d.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(s in exported for s in declared_symbols())


def test_config_struct_layout_matches():
    c = m2d.Config()
    for f, _ in m2d.Config._fields_:
        setattr(c, f, 7)
    m2d.lib().m2d_config_default(C.byref(c))
    assert c.scale == 1.0 and c.band_number == 5 and c.shard_count == 1 and c.resolution == 0.0
    assert c.weight_type == 0 and c.force_float == 0 and c.collect_stats == 0 and c.batch_frames == 0
    # header struct: 2 doubles + 15 ints (+ padding to 8)
    assert C.sizeof(m2d.Config) == 2 * 8 + 16 * 4
    assert c.render_blend == 0 and c.render_bands == 0 and c.f32_mode == 0
    assert C.sizeof(m2d.Stats) == 8 * (3 + 3 * m2d.MAX_LEVELS + 1 + 2 * m2d.MAX_LEVELS)  # ... + need_px, needw_px[MAX_LEVELS]


def test_unsupported_types_and_arguments():
    L = m2d.lib()
    h = C.c_void_p()
    cfg = m2d.default_config()
    assert L.m2d_create(0, C.byref(cfg), C.byref(h)) == -5 and not h.value      # NoType
    bad = m2d.default_config(render_blend=3)
    assert L.m2d_create(4, C.byref(bad), C.byref(h)) == -1 and not h.value      # TypeRender: unknown blend
    assert L.m2d_create(5, C.byref(cfg), C.byref(h)) == -5 and not h.value      # no such Map2DType
    ff = m2d.default_config(force_float=1)
    assert L.m2d_create(3, C.byref(ff), C.byref(h)) == -5 and not h.value       # ForceFloat unsupported
    assert m2d.Map2D.create(0) is None                                           # null SPtr in the reference
    assert L.m2d_sync(None) == -1 and L.m2d_queue_size(None) == 0


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    cfg = m2d.default_config()
    assert m2d.lib().m2d_create(1, C.byref(cfg), C.byref(h)) == -3 and not h.value  # M2D_ERR_CUDA, loudly
    with pytest.raises(RuntimeError):
        m2d.Map2D.create(3, thread=False)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pi-slam-fusion_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "map2d_oracle" not in txt, fn


def test_map2d_update_gps_corners_follow_the_reference_formula():
    """m2d_tile_gps_corners (pure host arithmetic, callable without a GPU) against a restatement of
    MultiBandMap2DCPU.cpp:709-757 + pi::calcLngLatFromDistance (PIL/src/hardware/Gps/utils_GPS.cpp:133-160)."""
    import math

    def qmul(l, r):   # SO3.h:435-442, (x, y, z, w)
        return (l[3] * r[0] + l[0] * r[3] + l[1] * r[2] - l[2] * r[1], l[3] * r[1] + l[1] * r[3] + l[2] * r[0] - l[0] * r[2],
                l[3] * r[2] + l[2] * r[3] + l[0] * r[1] - l[1] * r[0], l[3] * r[3] - l[0] * r[0] - l[1] * r[1] - l[2] * r[2])

    def expected(plane, gmin, ele, tx, ty, org):
        f32 = lambda v: float(np.float32(v))
        x0, y0 = f32(gmin[0] + tx * ele), f32(gmin[1] + ty * ele)
        x1, y1 = f32(x0 + ele), f32(y0 + ele)
        a, d2r, f = 6378137.0, 0.017453292519943, 1.0 / 298.257223563
        e2 = 2 * f - f * f
        phi = org[1] * d2r
        lng_unit = d2r * a * math.cos(phi) / math.sqrt(1 - e2 * math.sin(phi) ** 2)
        lat_unit = d2r * a * (1 - e2) / math.pow(1 - e2 * math.sin(phi) ** 2, 1.5)
        q = tuple(plane[3:7])
        out = []
        for (x, y) in ((x0, y0), (x1, y1)):
            r = qmul(qmul(q, (x, y, 0.0, 0.0)), (-q[0], -q[1], -q[2], q[3]))
            out.append((( plane[0] + r[0]) / lng_unit + org[0], (plane[1] + r[1]) / lat_unit + org[1], 0.0))
        return out

    rng = np.random.default_rng(3)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        plane = np.concatenate([rng.uniform(-500, 500, 3), q])
        grid = {"min": np.array([rng.uniform(-3000, 0), rng.uniform(-3000, 0), 0.0]), "length_pixel": rng.uniform(0.02, 0.5)}
        org = (rng.uniform(-180, 180), rng.uniform(-80, 80))
        tx, ty = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        tl, br = m2d.tile_gps_corners(plane, grid, tx, ty, org)
        etl, ebr = expected(plane, grid["min"], 256.0 * grid["length_pixel"], tx, ty, org)
        assert np.allclose(tl, etl, rtol=0, atol=1e-12) and np.allclose(br, ebr, rtol=0, atol=1e-12)
    # identity plane, tile (0,0) at the origin: one tile of 256 * 0.1 m east and north of the GPS origin
    grid = {"min": np.zeros(3), "length_pixel": 0.1}
    tl, br = m2d.tile_gps_corners([0, 0, 0, 0, 0, 0, 1], grid, 0, 0, (108.0, 34.0))
    assert tuple(tl) == (108.0, 34.0, 0.0)
    assert abs((br[0] - 108.0) * 111320 * math.cos(math.radians(34)) - 25.6) < 0.1 and abs((br[1] - 34.0) * 110922 - 25.6) < 0.1


def test_map2d_update_command_and_overlay_pixmap():
    """The rest of the Map2DUpdate hand-over, on the CPU stand-in: the command string format of
    MultiBandMap2DCPU.cpp:754-755 and the pixmap Map2DItemHandle builds (Map2DItem.cpp:56-84)."""
    from oracle import oracle as O
    import pi_slam_fusion_b200.synth as synth
    grid = {"min": np.zeros(3), "length_pixel": 0.1}
    cmd = m2d.map2d_update_command([0, 0, 0, 0, 0, 0, 1], grid, 0, 0, (108.888931, 34.257287, 400.0))
    tok = cmd.split()
    assert tok[:2] == ["Map2DUpdate", "LastTexMat"] and len(tok) == 8
    assert tok[2] == "108.888931000" and tok[3] == "34.257287000" and tok[4] == "0.000000000" and tok[7] == "0.000000000"
    assert all(re.fullmatch(r"-?\d+\.\d{9}", t) for t in tok[2:])
    seq = synth.Sequence(3, 320, 180, seed=5, fpl=3, prepare_frames=3)
    for typ in (1, 3):
        o = O.OracleMap2D.create(typ)
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(seq.n):
            assert o.feed(seq.frame(k), seq.poses[k])
        x0, y0, x1, y1 = o.last_rect()
        pix = m2d.tile_overlay(o, x0, y0, False)
        img = o.get_tile_image(x0, y0, False)
        if typ == 3:
            w0 = o.get_tile(x0, y0, 0)[1]
            assert pix.shape == (256, 256, 4) and np.array_equal(pix[::-1, :, :3], img)
            assert np.array_equal(pix[::-1, :, 3] == 255, w0 != 0) and set(np.unique(pix[..., 3])) <= {0, 255}
        else:
            assert pix.shape == (256, 256, 3) and np.array_equal(pix[::-1, :, ::-1], img[..., :3])


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/map2d_b200.h is a C header (C99, -pedantic clean): a C program links against the library, gets a loud
    failure from m2d_create on a box without a GPU, and can call the handle-free host helpers."""
    exe = str(tmp_path / "cabi_c")
    libdir = os.path.dirname(m2d.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cabi_c", "main.c"), "-o", exe, "-L", libdir, "-lmap2d_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    lng, lat = (float(v) for v in lines[-1].split())
    assert 108.0002 < lng < 108.0004 and 34.0002 < lat < 34.0003
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert "rc=-3" in lines[0] and "no CPU path" in out.stderr
