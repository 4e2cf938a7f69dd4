"""Oracle (CPU, no cv2 needed) and CUDA path (gpu) against the committed golden fixtures of tests/golden/
(generated from real OpenCV 4.13 by tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
import pi_slam_fusion_b200.synth as synth

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLD = json.load(f)
KAT = np.load(os.path.join(HERE, "golden", "kat.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def state_record(m, typ, levels=6):
    rec = {}
    g = m.grid()
    for ty in range(g["h"]):
        for tx in range(g["w"]):
            if typ == 1:
                t = m.get_tile(tx, ty)
                if t is not None:
                    rec["%d,%d" % (tx, ty)] = sha(t)
            else:
                if m.get_tile(tx, ty, 0) is None:
                    continue
                rec["%d,%d" % (tx, ty)] = [[sha(m.get_tile(tx, ty, l)[0]), sha(m.get_tile(tx, ty, l)[1])] for l in range(levels)]
    return rec


@pytest.fixture(autouse=True)
def _mode():
    O.set_f32_mode(0)
    O.set_threads(1)
    yield
    O.set_f32_mode(0)


def test_primitive_kats():
    M = KAT["H"]
    assert np.array_equal(O.get_perspective_transform(KAT["H_src"], KAT["H_dst"]), M)
    assert np.array_equal(O.invert3x3(M), KAT["H_inv"])
    p = GOLD["cv2"]["primitives"]
    assert sha(O.warp_u8c4(KAT["warp_src_u8c4"], M, (256, 256))) == p["warp_u8c4_256"]
    assert sha(O.warp_s16c3_reflect(KAT["warp_src_s16c3"], M, (256, 256))) == p["warp_s16c3_reflect_256"]
    assert sha(O.warp_f32_nearest(KAT["warp_src_f32"], M, (256, 256))) == p["warp_f32_nearest_249_256"]
    for key in ("16x16", "5x7", "64x48"):
        a, f = KAT["pyr_s16_" + key], KAT["pyr_f32_" + key]
        assert np.array_equal(O.pyrdown_s16(a), KAT["pyrdown_s16_" + key])
        assert np.array_equal(O.pyrup_s16(a), KAT["pyrup_s16_" + key])
        assert np.array_equal(O.pyrdown_f32(f), KAT["pyrdown_f32_249_" + key])
        O.set_f32_mode(1)
        assert np.array_equal(O.pyrdown_f32(f), KAT["pyrdown_f32_cv2_" + key])
        O.set_f32_mode(0)


@pytest.mark.parametrize("name", sorted(GOLD["sequences"]))
@pytest.mark.parametrize("typ", [1, 3])
def test_oracle_reproduces_opencv_goldens(name, typ):
    gold = GOLD["cv2"]["%s/type%d" % (name, typ)]
    seq = synth.Sequence(**GOLD["sequences"][name])
    O.set_f32_mode(1)  # cv2 4.x float association for the weight pyramid; everything else is mode-independent
    o = O.OracleMap2D(typ)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(seq.n):
        assert o.feed(seq.frame(k), seq.poses[k]) == gold["accepted"][k]
        if gold["accepted"][k]:
            assert list(o.last_rect()) == gold["rects"][k]
    g = o.grid()
    assert (g["w"], g["h"]) == (gold["grid"]["w"], gold["grid"]["h"])
    assert list(g["min"]) == gold["grid"]["min"] and list(g["max"]) == gold["grid"]["max"]
    assert g["length_pixel"] == gold["grid"]["length_pixel"]
    assert state_record(o, typ) == gold["tiles"]
    img, org = o.get_image()
    assert sha(img) == gold["mosaic"] and list(org) == gold["mosaic_origin"] and list(img.shape) == gold["mosaic_shape"]


@pytest.mark.parametrize("name", sorted(GOLD["sequences"]))
@pytest.mark.parametrize("typ", [1, 3])
def test_oracle_default_mode_goldens(name, typ):
    gold = GOLD["oracle249"]["%s/type%d" % (name, typ)]
    seq = synth.Sequence(**GOLD["sequences"][name])
    o = O.OracleMap2D(typ)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(seq.n):
        o.feed(seq.frame(k), seq.poses[k])
    assert state_record(o, typ) == gold["tiles"]
    assert sha(o.get_image()[0]) == gold["mosaic"]
    st = o.stats()   # (need_px is a GPU-pipeline counter added later: always 0 from the oracle, not part of the goldens)
    assert {k: st[k] for k in gold["stats"]} == gold["stats"] and not any(st["need_px"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["sequences"]))
@pytest.mark.parametrize("typ", [1, 3])
def test_cuda_path_reproduces_goldens(name, typ):
    """The CUDA path through the C-ABI against the committed fixtures: tile rects and grid against the OpenCV
    goldens (float-association independent), raw state and mosaic against the 2.4.9-association goldens."""
    import pi_slam_fusion_b200.map2d as m2d
    cvg = GOLD["cv2"]["%s/type%d" % (name, typ)]
    gold = GOLD["oracle249"]["%s/type%d" % (name, typ)]
    seq = synth.Sequence(**GOLD["sequences"][name])
    g = m2d.Map2D.create(typ, thread=False, collect_stats=1)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(seq.n):
        assert g.feed(seq.frame(k), seq.poses[k]) == cvg["accepted"][k]
        if cvg["accepted"][k]:
            assert list(g.last_rect()) == cvg["rects"][k]
    gg = g.grid()
    assert (gg["w"], gg["h"]) == (cvg["grid"]["w"], cvg["grid"]["h"])
    assert list(gg["min"]) == cvg["grid"]["min"] and list(gg["max"]) == cvg["grid"]["max"]
    assert state_record(g, typ) == gold["tiles"]
    img, org = g.get_image()
    assert sha(img) == gold["mosaic"] and list(org) == gold["mosaic_origin"]
    if typ == 1:  # weighted state is integer-only: it must also equal the pure-OpenCV golden
        assert state_record(g, typ) == cvg["tiles"] and sha(img) == cvg["mosaic"]
    s = g.stats()
    for k in ("frames_fused", "input_px", "region_px", "fresh_px", "win_px"):
        assert s[k] == gold["stats"][k], k
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", ["weights_first", "dense", "streaming"])
@pytest.mark.parametrize("jitter", [False, True])
def test_gpu_multiband_equals_real_opencv_end_to_end(monkeypatch, pipeline, jitter):
    """The CUDA multi-band path against REAL OpenCV, with nothing of ours in between: tests/cv2_reference.py builds the whole
    feed()/save() recipe from cv2 4.13 primitives (getPerspectiveTransform, warpPerspective 16SC3 BORDER_REFLECT, pyrDown f32 and
    16S, detail.createLaplacePyr / restoreImageFromLaplacePyr; only the nearest weight warp is restated, because cv2 >= 4.12
    changed its border rule).  With m2d_config.f32_mode = 1 (cv2 4.x's float association of the weight pyrDown instead of
    2.4.9's) the GPU must reproduce its grid, every tile of every level -- int16 Laplacians AND float weights -- and the
    saved mosaic byte for byte, through the weights-first pipeline (culling, work lists, TMA pyrDown), the dense pipeline
    and frame-by-frame streaming."""
    import torch
    import pi_slam_fusion_b200.map2d as m2d
    pytest.importorskip("cv2")
    from tests.cv2_reference import Cv2Map2D
    if pipeline == "dense":
        monkeypatch.setenv("M2D_SPARSE", "0")
    seq = synth.Sequence(14, 320, 180, seed=7, jitter=jitter, noise=jitter, fpl=4, prepare_frames=6)
    frames = seq.frames()
    c = Cv2Map2D(3)
    g = m2d.Map2D.create(3, thread=False, f32_mode=1, batch_frames=0 if pipeline == "weights_first" else 5)
    assert c.prepare(seq.plane, seq.camera, seq.prepare_poses) and g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    exp = [c.feed(frames[k], seq.poses[k]) for k in range(seq.n)]
    if pipeline == "streaming":
        got = [g.feed(frames[k], seq.poses[k]) for k in range(seq.n)]
    else:
        dev = torch.from_numpy(frames).cuda()
        got = [r == 0 for r in g.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, True)]
    assert got == exp
    g.sync()
    gr = g.grid()
    assert (gr["w"], gr["h"]) == (c.w, c.h) and gr["length_pixel"] == c.lp and list(gr["min"]) == c.min
    assert len(c.tiles) > 0 and g.tile_count() == len(c.tiles)
    for (tx, ty) in c.tiles:
        for l in range(6):
            gl, gw = g.get_tile(tx, ty, l)
            cl, cw = c.get_tile(tx, ty, l)
            assert np.array_equal(gw, cw), "tile (%d,%d) L%d weights differ from cv2" % (tx, ty, l)
            assert np.array_equal(gl, cl), "tile (%d,%d) L%d Laplacian differs from cv2" % (tx, ty, l)
    (ig, og), (ic, oc) = g.get_image(), c.get_image()
    assert og == oc and np.array_equal(ig, ic)
    g.close()


# ---- Map2DRender (type 4) ---------------------------------------------------------------------------------------------
RENDER_KEYS = sorted(GOLD.get("render", {}))


def _render_record(m, res):
    r16, mask, nb, org = m.render_get()
    return {"accepted": [int(v) for v in res], "bands": nb, "origin": list(org), "shape": list(r16.shape), "result": sha(r16),
            "mask": sha(mask), "image8": sha(np.clip(r16, 0, 255).astype(np.uint8))}


@pytest.mark.parametrize("key", RENDER_KEYS)
def test_render_oracle_reproduces_goldens(key):
    """tests/golden/make_golden.py recorded these only after the weighted-sum blends had been found equal to the real
    cv2.detail_MultiBandBlender; no cv2 needed here."""
    name, blend, fam = key.split("/")
    seq = synth.Sequence(**GOLD["sequences"][name])
    O.set_f32_mode(1 if fam == "cv2" else 0)
    o = O.OracleMap2D(O.TYPE_RENDER, render_blend=int(blend[-1]))
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    rc, res = o.render_frames(seq.frames(), seq.poses)
    assert rc == 0 and _render_record(o, res) == GOLD["render"][key]


@pytest.mark.gpu
@pytest.mark.parametrize("key", RENDER_KEYS)
def test_render_cuda_path_reproduces_goldens(key):
    """m2d_render_frames through the C-ABI against the committed fixtures, in both float associations (f32_mode 1 = the
    fixtures that equal the real OpenCV blender; 0 = the 2.4.9 association the library defaults to)."""
    import pi_slam_fusion_b200.map2d as m2d
    name, blend, fam = key.split("/")
    seq = synth.Sequence(**GOLD["sequences"][name])
    g = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=False, render_blend=int(blend[-1]), f32_mode=1 if fam == "cv2" else 0)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    rc, res = g.render_frames(seq.frames(), seq.poses)
    assert rc == 0
    rec = _render_record(g, res)
    assert rec == GOLD["render"][key]
    img, _ = g.get_image()
    assert sha(img) == rec["image8"]
    g.close()
