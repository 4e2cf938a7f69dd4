"""Map2DRender (Map2D type 4) on the GPU, through the C-ABI, against the CPU oracle -- whose blends are
pinned to real OpenCV in tests/test_render_oracle.py.  Bar: bit-exact (the blender's raw CV_16SC3 result, its mask, the band
count, the canvas placement, the per-frame accept list and the 8-bit image m2d_get_image returns)."""
import os

import numpy as np
import pytest

import pi_slam_fusion_b200.map2d as m2d
import pi_slam_fusion_b200.synth as synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_mode():
    O.set_f32_mode(0)
    O.set_threads(max(1, (os.cpu_count() or 2) // 2))
    yield
    O.set_f32_mode(0)
    O.set_threads(1)


def run_pair(seq, frames=None, poses=None, on_device=False, **cfg):
    frames = seq.frames() if frames is None else frames
    poses = seq.poses if poses is None else poses
    O.set_f32_mode(cfg.get("f32_mode", 0))
    g = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=False, **cfg)
    o = O.OracleMap2D(O.TYPE_RENDER, **cfg)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    if on_device:
        import torch
        dev = torch.from_numpy(frames).cuda()
        rc_g, res_g = g.render_frames(dev.data_ptr(), poses, on_device=True, w=seq.w, h=seq.h)
        g.sync()
    else:
        rc_g, res_g = g.render_frames(frames, poses)
    rc_o, res_o = o.render_frames(frames, poses)
    assert rc_g == rc_o and np.array_equal(res_g, res_o)
    return g, o, res_o


def compare(g, o):
    a, b = g.render_get(), o.render_get()
    assert (a is None) == (b is None)
    if a is None:
        return None
    assert a[2] == b[2], "band count %d vs %d" % (a[2], b[2])
    assert a[3] == b[3], "canvas origin"
    assert a[0].shape == b[0].shape
    assert np.array_equal(a[1], b[1]), "mask: %d px differ" % int((a[1] != b[1]).sum())
    assert np.array_equal(a[0], b[0]), "result: %d values differ" % int((a[0] != b[0]).sum())
    og, oo = g.grid(), o.grid()
    assert (og["w"], og["h"]) == (oo["w"], oo["h"]) and np.array_equal(og["min"], oo["min"])
    img, org = g.get_image()
    assert np.array_equal(img, np.clip(b[0], 0, 255).astype(np.uint8))
    assert a[1].any()
    return a


@pytest.mark.parametrize("blend", [0, 1, 2])
@pytest.mark.parametrize("bands", [0, 3, 6])
def test_render_equals_oracle(blend, bands):
    seq = synth.Sequence(10, 320, 180, seed=21 + blend, jitter=True, fpl=4, prepare_frames=5)
    g, o, _ = run_pair(seq, render_blend=blend, render_bands=bands)
    r = compare(g, o)
    assert bands == 0 or r[2] == bands
    g.close()


@pytest.mark.parametrize("blend,f32_mode", [(0, 1), (1, 1), (0, 0)])
def test_render_720p_nadir_and_noise(blend, f32_mode):
    """Exactly nadir frames (affine homographies: the constant-denominator fast path of the coordinate code) with i.i.d. noise
    content, device-resident frames, both float associations."""
    seq = synth.Sequence(12, 1280, 720, seed=5, jitter=False, noise=True, prepare_frames=6)
    g, o, _ = run_pair(seq, on_device=True, render_blend=blend, f32_mode=f32_mode)
    compare(g, o)
    g.close()


def test_render_spreads_the_map_and_skips_oblique_frames():
    seq = synth.Sequence(14, 320, 180, seed=8, jitter=True, fpl=14, along=0.8, prepare_frames=3)
    seq.prepare_poses = seq.poses[-3:]     # the prepared grid covers the END of the flight line: the batch grows it towards -y
    poses = seq.poses.copy()
    # frame 5: pitched 70 degrees -> a corner ray misses the 0.4 test -> skipped (Map2DRender.cpp:548-557)
    poses[5, 3:] = synth._qmul(synth._qaxis((0, 1, 0), np.radians(70.0)), np.array([1.0, 0.0, 0.0, 0.0]))
    g, o, res = run_pair(seq, poses=poses)
    assert res[5] == 1 and (np.delete(res, 5) == 0).all()
    gh = o.grid()["h"]
    r = compare(g, o)
    o2 = O.OracleMap2D(O.TYPE_RENDER)
    assert o2.prepare(seq.plane, seq.camera, seq.prepare_poses) and o2.grid()["h"] < gh and r[3][1] < 0   # the batch did spread the map
    g.close()


def test_render_chunks_do_not_change_results(monkeypatch):
    """A batch larger than the scratch budget is blended chunk by chunk, in feed order: same bits."""
    monkeypatch.setenv("M2D_SCRATCH_GB", "0.25")
    seq = synth.Sequence(24, 1280, 720, seed=13, jitter=True, prepare_frames=8)
    g, o, _ = run_pair(seq, render_blend=1)
    compare(g, o)
    assert g.launch_count() >= 2 * (1 + g.render_get()[2] + 1)   # at least two chunks of warp + pyramid + blend
    g.close()


def test_render_weighted_sum_on_gpu_equals_cv2_blender():
    """GPU (f32_mode 1 = cv2 4.x float association) against the REAL cv2.detail_MultiBandBlender, fed with the warped frames
    of the oracle -- whose warps are themselves pinned to cv2.warpPerspective."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    seq = synth.Sequence(9, 320, 180, seed=2, jitter=True, fpl=3, prepare_frames=4)
    for blend, wt in ((1, cv2.CV_32F), (2, cv2.CV_16S)):
        g, o, _ = run_pair(seq, render_blend=blend, f32_mode=1)
        res, mask, nb, _ = g.render_get()
        b = cv2.detail_MultiBandBlender(0, nb, wt)
        b.prepare((0, 0, res.shape[1], res.shape[0]))
        for i in range(seq.n):
            img, m, c = o.render_warped(i)
            b.feed(img, m, c)
        ref, ref_mask = b.blend(None, None)
        assert np.array_equal(ref_mask, mask) and np.array_equal(ref, res)
        g.close()


def test_render_boundary_semantics():
    """feed() on a TypeRender map: thread=False returns false (Map2DRender::renderFrame, :464-467); thread=True queues only;
    prepare(thread=True) renders the prepare-frames as the one batch; a second batch replaces the canvas; reset drops it."""
    seq = synth.Sequence(8, 320, 180, seed=4, jitter=True, fpl=4, prepare_frames=6)
    frames = seq.frames()
    g = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=False)
    assert g.render_get() is None
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert g.feed(frames[0], seq.poses[0]) is False
    assert g.render_get() is None and g.get_image() is None
    g.close()
    t = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=True)
    assert t.prepare(seq.plane, seq.camera, [(frames[k], seq.poses[k]) for k in range(6)])
    assert t.feed(frames[6], seq.poses[6]) is True
    o = O.OracleMap2D(O.TYPE_RENDER)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    o.render_frames(frames[:6], seq.poses[:6])
    compare(t, o)
    t.render_frames(frames[2:], seq.poses[2:])
    o.render_frames(frames[2:], seq.poses[2:])
    compare(t, o)
    bad = np.zeros((2, 90, 160, 3), np.uint8)
    rc, _ = t.render_frames(bad, seq.poses[:2])
    assert rc == 1                                  # frame size != camera size: refused like renderFrame
    t.close()


def test_render_edge_cases():
    """A batch of oblique frames only: the reference blends nothing onto the one tile around the origin and returns true (an
    all-masked canvas); an empty batch is an argument error."""
    seq = synth.Sequence(4, 320, 180, seed=3, jitter=True, fpl=3, prepare_frames=4)
    frames = seq.frames()
    g, o, _ = run_pair(seq)
    poses = seq.poses.copy()
    for k in range(seq.n):
        poses[k, 3:] = synth._qmul(synth._qaxis((0, 1, 0), np.radians(75.0)), np.array([1.0, 0.0, 0.0, 0.0]))
    rc_g, res_g = g.render_frames(frames, poses)
    rc_o, res_o = o.render_frames(frames, poses)
    assert rc_g == rc_o == 0 and np.array_equal(res_g, res_o) and (res_g == 1).all()
    a, b = g.render_get(), o.render_get()
    assert a[2] == b[2] and a[3] == b[3] and a[0].shape == b[0].shape == (256, 256, 3)
    assert not a[0].any() and not a[1].any() and not b[0].any() and not b[1].any()
    with pytest.raises(RuntimeError):
        g.render_frames(frames[:0], poses[:0])
    g.close()
