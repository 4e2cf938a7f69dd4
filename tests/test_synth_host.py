"""Host-side logic that needs no GPU: synthetic generator determinism/geometry and the oracle's edge cases
(rejects, spreadMap, re-prepare) that the GPU tests mirror."""
import numpy as np

from oracle import oracle as O
import pi_slam_fusion_b200.synth as synth


def test_sequence_is_deterministic_and_nadir():
    a = synth.Sequence(30, 320, 180, seed=5, jitter=True)
    b = synth.Sequence(30, 320, 180, seed=5, jitter=True)
    assert np.array_equal(a.poses, b.poses) and np.array_equal(a.frame(7), b.frame(7))
    c = synth.Sequence(30, 320, 180, seed=5)
    assert np.allclose(c.poses[:, 3:], [1, 0, 0, 0]) and np.allclose(c.poses[:, 2], 100.0)
    assert a.frame(3).shape == (180, 320, 3) and a.frame(3).dtype == np.uint8
    # 80 % along-track overlap, 60 % cross-track
    fpl = synth.frames_per_line(30, 320, 180)
    step = c.poses[1, 1] - c.poses[0, 1]
    assert abs(step / c.gsd - 0.2 * 180) < 1.5 and fpl >= 2


def test_footprint_equals_frame_size_at_scale_one():
    seq = synth.Sequence(4, 1280, 720, seed=1)
    o = O.OracleMap2D(1)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert abs(o.grid()["length_pixel"] - seq.gsd) < 1e-12
    assert o.feed(seq.frame(0), seq.poses[0])
    x0, y0, x1, y1 = o.last_rect()
    assert (x1 - x0, y1 - y0) in ((6, 4), (6, 3), (5, 4), (5, 3), (6, 5))


def test_prepare_rejections():
    seq = synth.Sequence(4, 320, 180, seed=1)
    o = O.OracleMap2D(1)
    assert not o.prepare(seq.plane, seq.camera, seq.poses[:0])                       # no frames (Map2D.cpp:35)
    bad = seq.camera.copy(); bad[2] = 0
    assert not o.prepare(seq.plane, bad, seq.poses)                                  # fx == 0
    straddle = seq.poses.copy(); straddle[1, 2] = -50
    assert not o.prepare(seq.plane, seq.camera, straddle)                            # heights straddle (Map2DCPU.cpp:62)
    assert not o.feed(seq.frame(0), seq.poses[0])                                    # not prepared (Map2DCPU.cpp:129)
    assert o.prepare(seq.plane, seq.camera, seq.poses)
    assert not o.feed(np.zeros((10, 10, 3), np.uint8), seq.poses[0])                 # wrong size (:158-162)
    oblique = seq.poses[0].copy(); oblique[3:] = [0.5, 0.5, 0.5, 0.5]
    assert not o.feed(seq.frame(0), oblique)                                         # oblique (:179-182)


def test_spread_map_keeps_tiles():
    seq = synth.Sequence(12, 256, 144, seed=3, jitter=True, fpl=3, prepare_frames=2, cross=1.5, along=0.9)
    for typ in (1, 3):
        o = O.OracleMap2D(typ)
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        g0 = o.grid()
        assert o.feed(seq.frame(0), seq.poses[0])
        before, org0 = o.get_image()
        for k in range(1, seq.n):
            assert o.feed(seq.frame(k), seq.poses[k])
        g1 = o.grid()
        assert g1["w"] * g1["h"] > g0["w"] * g0["h"]                                # grid grew, never shrinks
        assert g1["min"][0] <= g0["min"][0] and g1["max"][0] >= g0["max"][0]
        # grid stays anchored on whole tiles
        assert abs(((g0["min"][0] - g1["min"][0]) / (256 * g1["length_pixel"])) - round((g0["min"][0] - g1["min"][0]) / (256 * g1["length_pixel"]))) < 1e-6


def test_multiband_levels_and_band_clamp():
    assert O.OracleMap2D(3, band_number=5).levels == 6
    assert O.OracleMap2D(3, band_number=20).levels == 9                              # clamp to log2(256)=8 bands
    seq = synth.Sequence(3, 256, 144, seed=2, fpl=3, prepare_frames=3)
    o = O.OracleMap2D(3, band_number=8)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert o.feed(seq.frame(0), seq.poses[0])
    x0, y0, _, _ = o.last_rect()
    lap, wgt = o.get_tile(x0, y0, 8)
    assert lap.shape == (1, 1, 3) and wgt.shape == (1, 1)
