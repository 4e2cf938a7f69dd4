"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded
inputs.  Bar: tile footprint / indexing and all integer state bit-exact; float weights bit-exact too (tolerance
stated: max relative error <= 1e-5 is the contract, we assert equality and report)."""
import os

import numpy as np
import pytest

import pi_slam_fusion_b200.map2d as m2d
import pi_slam_fusion_b200.synth as synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

W_TOL = 1e-5  # north_star: float state max relative error


def run_pair(typ, seq, **cfg):
    g = m2d.Map2D.create(typ, thread=False, **cfg)
    o = O.OracleMap2D.create(typ, **cfg)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) == o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    go, oo = g.grid(), o.grid()
    assert (go["w"], go["h"]) == (oo["w"], oo["h"])
    assert np.array_equal(go["min"], oo["min"]) and np.array_equal(go["max"], oo["max"]) and go["length_pixel"] == oo["length_pixel"]
    for k in range(seq.n):
        f = seq.frame(k)
        rg, ro = g.feed(f, seq.poses[k]), o.feed(f, seq.poses[k])
        assert rg == ro, "frame %d accept mismatch" % k
        if rg:
            assert g.last_rect() == o.last_rect(), "frame %d tile rect mismatch" % k
    g.sync()
    return g, o


def compare_state(g, o, typ):
    gg, og = g.grid(), o.grid()
    assert (gg["w"], gg["h"]) == (og["w"], og["h"])
    assert np.array_equal(gg["min"], og["min"]) and np.array_equal(gg["max"], og["max"])
    ntiles = 0
    for ty in range(og["h"]):
        for tx in range(og["w"]):
            if typ == 1:
                ot, gt = o.get_tile(tx, ty), g.get_tile(tx, ty)
                assert (ot is None) == (gt is None), "tile (%d,%d) presence" % (tx, ty)
                if ot is None:
                    continue
                ntiles += 1
                assert np.array_equal(ot, gt), "tile (%d,%d): %d bytes differ" % (tx, ty, int((ot != gt).sum()))
            else:
                for l in range(o.levels):
                    ot, gt = o.get_tile(tx, ty, l), g.get_tile(tx, ty, l)
                    assert (ot is None) == (gt is None), "tile (%d,%d) presence" % (tx, ty)
                    if ot is None:
                        break
                    ntiles += (l == 0)
                    assert np.array_equal(ot[1], gt[1]), "tile (%d,%d) L%d weights: max rel err %g" % (
                        tx, ty, l, float(np.max(np.abs(ot[1] - gt[1]) / np.maximum(np.abs(ot[1]), 1e-30))))
                    assert np.array_equal(ot[0], gt[0]), "tile (%d,%d) L%d laplacian: %d values differ" % (
                        tx, ty, l, int((ot[0] != gt[0]).sum()))
    assert ntiles > 0
    gi, oi = g.get_image(), o.get_image()
    assert gi[1] == oi[1] and gi[0].shape == oi[0].shape
    assert np.array_equal(gi[0], oi[0]), "mosaic: %d bytes differ" % int((gi[0] != oi[0]).sum())
    return ntiles


@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("jitter,noise", [(False, False), (True, True)])
def test_small_sequence_bit_exact(typ, jitter, noise):
    seq = synth.Sequence(14, 320, 180, seed=7, jitter=jitter, noise=noise, fpl=5, prepare_frames=6)
    g, o = run_pair(typ, seq)
    compare_state(g, o, typ)
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_spread_map_and_rejects(typ):
    """Prepare on 3 frames only so later frames leave the grid (spreadMap, Map2DCPU.cpp:339-382); include an
    oblique pose that must be rejected (Map2DCPU.cpp:179-182) and a wrong-sized frame (:158-162)."""
    seq = synth.Sequence(12, 256, 144, seed=3, jitter=True, fpl=3, prepare_frames=2, cross=1.5, along=0.9)
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    w0 = g.grid()["w"] * g.grid()["h"]
    for k in range(seq.n):
        f = seq.frame(k)
        assert g.feed(f, seq.poses[k]) == o.feed(f, seq.poses[k])
        assert g.last_rect() == o.last_rect()
    oblique = seq.poses[0].copy()
    oblique[3:] = [0.5, 0.5, 0.5, 0.5]  # optical axis horizontal
    assert g.feed(seq.frame(0), oblique) is False and o.feed(seq.frame(0), oblique) is False
    assert g.feed(np.zeros((10, 10, 3), np.uint8), seq.poses[0]) is False
    g.sync()
    assert g.grid()["w"] * g.grid()["h"] > w0, "test did not exercise spreadMap"
    compare_state(g, o, typ)
    g.close()


def test_weight_type_and_scale():
    seq = synth.Sequence(8, 320, 180, seed=5, jitter=True, fpl=4, prepare_frames=4)
    for typ in (1, 3):
        g, o = run_pair(typ, seq, weight_type=1, scale=1.5, background=255)
        compare_state(g, o, typ)
        g.close()


def test_band_number_variants():
    seq = synth.Sequence(6, 320, 180, seed=9, jitter=True, fpl=3, prepare_frames=3)
    for bands in (1, 3, 8):
        g, o = run_pair(3, seq, band_number=bands)
        assert g.levels == o.levels == bands + 1
        compare_state(g, o, 3)
        g.close()


def test_bounds_kernel_matches_oracle():
    seq = synth.Sequence(64, 1280, 720, seed=2, jitter=True)
    g = m2d.Map2D.create(1, thread=False)
    o = O.OracleMap2D.create(1)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    poses = seq.poses.copy()
    poses[5, 3:] = [0.5, 0.5, 0.5, 0.5]
    rg, hg = g.compute_bounds(poses)
    ro, ho = o.compute_bounds(poses)
    assert np.array_equal(rg, ro)
    assert np.array_equal(hg, ho), "inverse homographies differ in %d entries" % int((hg != ho).sum())
    assert (rg[5] == -1).all()
    g.close()


def test_plan_rects_and_pose_only_feed():
    """m2d_plan_rects predicts the absolute rect of every frame exactly (spreadMap included) without touching the
    map; m2d_feed_poses grows the grid like feed() and refuses poses under which the shard owns tiles."""
    seq = synth.Sequence(40, 320, 180, seed=9, jitter=True, fpl=4, prepare_frames=3, cross=0.9, along=0.6)
    poses = seq.poses.copy()
    poses[7, 3:] = [0.5, 0.5, 0.5, 0.5]
    g = m2d.Map2D.create(1, thread=False)
    o = O.OracleMap2D.create(1)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    g0 = g.grid()
    plan = g.plan_rects(poses)
    assert np.array_equal(plan, o.plan_rects(poses)) and (plan[7] == -1).all()
    g1 = g.grid()
    assert (g0["w"], g0["h"]) == (g1["w"], g1["h"]) and g.tile_count() == 0, "the dry run must not touch the map"
    # a shard that owns nothing under the survey: pose-only feed of everything reproduces the unsharded grid
    far = int(plan[plan[:, 2] > 0, 2].max()) + 5
    g.set_shard(1, 2, 0, 1 << 20, far - (1 << 20))   # rank 1 owns abs x >= far only
    res = g.feed_poses(poses)
    frames = seq.frames()
    exp = [0 if o.feed(frames[k], poses[k]) else 1 for k in range(seq.n)]
    assert res.tolist() == exp and g.tile_count() == 0
    gg, go = g.grid(), o.grid()
    assert (gg["w"], gg["h"]) == (go["w"], go["h"]) and np.array_equal(gg["min"], go["min"]) and gg["w"] > g0["w"]
    # ... and one that owns everything is refused (its pixels are required), loudly
    g.reset()
    g.set_shard(0, 2, 0, 1 << 20, -(1 << 19))
    with pytest.raises(RuntimeError):
        g.feed_poses(poses[:3])
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_stats_match_oracle(typ):
    seq = synth.Sequence(10, 320, 180, seed=4, jitter=True, fpl=5, prepare_frames=5)
    g, o = run_pair(typ, seq, collect_stats=1)
    sg, so = g.stats(), o.stats()
    for k in ("frames_fed", "frames_fused", "input_px", "region_px", "fresh_px", "win_px"):
        assert sg[k] == so[k], (k, sg[k], so[k])
    if typ == 1:
        assert sg["footprint_px"] == so["footprint_px"]
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_full_size_720p_prefix(typ):
    """BASELINE.json configs[0]/[1] frame size, first 24 frames of the seeded survey, bit-exact vs the oracle."""
    seq = synth.Sequence(100 if typ == 1 else 500, 1280, 720, seed=typ)
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    O.set_threads(8)
    try:
        assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(24):
            f = seq.frame(k)
            assert g.feed(f, seq.poses[k]) == o.feed(f, seq.poses[k])
            assert g.last_rect() == o.last_rect()
        g.sync()
        compare_state(g, o, typ)
    finally:
        O.set_threads(1)
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_full_size_4000x3000_prefix(typ):
    """BASELINE.json configs[2]/[3] frame size (Phantom/Mavic-class 12 Mpx): a 17x13-tile region per frame, 200 MB of
    scratch pyramid per frame.  First frames of the seeded survey, fed as one batch, bit-exact vs the oracle."""
    import torch
    seq = synth.Sequence(1000, 4000, 3000, seed=3, jitter=(typ == 3))
    n = 3
    frames = np.stack([seq.frame(k) for k in range(n)])
    dev = torch.from_numpy(frames).cuda()
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    O.set_threads(16)
    try:
        assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        res = g.feed_batch(dev.data_ptr(), n, 4000 * 3000 * 3, 4000, 3000, 4000 * 3, seq.poses[:n], True)
        for k in range(n):
            assert (res[k] == 0) == o.feed(frames[k], seq.poses[k])
        g.sync()
        assert g.last_rect() == o.last_rect()
        compare_state(g, o, typ)
    finally:
        O.set_threads(1)
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_ragged_shapes_and_strides(typ):
    """Frame width not a multiple of 4 (generic pack path, unaligned rows), odd height, a padded row stride on the
    host path, and a 1-frame batch: everything must still match the oracle bit for bit."""
    import torch
    seq = synth.Sequence(7, 322, 181, seed=17, jitter=True, fpl=4, prepare_frames=3)
    g = m2d.Map2D.create(typ, thread=False, batch_frames=3)
    o = O.OracleMap2D.create(typ)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    padded = np.zeros((181, 322 * 3 + 13), np.uint8)
    for k in range(seq.n):
        f = seq.frame(k)
        assert o.feed(f, seq.poses[k])
        if k % 3 == 0:    # host image with a padded stride
            padded[:, :322 * 3] = f.reshape(181, -1)
            view = padded[:, :322 * 3].reshape(181, 322, 3)
            assert g.feed(view, seq.poses[k])
        elif k % 3 == 1:  # device image, tight
            d = torch.from_numpy(f).cuda()
            assert g.feed_device(d.data_ptr(), 322, 181, 322 * 3, seq.poses[k])
            g.sync()
        else:             # device image with a padded stride, as a 1-frame batch
            d = torch.from_numpy(padded.copy()).cuda()
            d[:, :322 * 3] = torch.from_numpy(f.reshape(181, -1)).cuda()
            res = g.feed_batch(d.data_ptr(), 1, 0, 322, 181, 322 * 3 + 13, seq.poses[k:k + 1], True)
            assert res[0] == 0
            g.sync()
    g.sync()
    compare_state(g, o, typ)
    assert g.feed_batch(0, 0, 0, 322, 181, 322 * 3, seq.poses[:0], True).size == 0  # empty batch is a no-op
    g.close()


def test_feed_paths_agree():
    """feed (host, staged), feed_device and feed_batch must leave identical state."""
    import torch
    seq = synth.Sequence(10, 320, 180, seed=8, jitter=True, fpl=5, prepare_frames=5)
    frames = seq.frames()
    dev = torch.from_numpy(frames).cuda()
    for typ in (1, 3):
        a = m2d.Map2D.create(typ, thread=False)
        b = m2d.Map2D.create(typ, thread=False)
        c = m2d.Map2D.create(typ, thread=False)
        for m in (a, b, c):
            assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(seq.n):
            assert a.feed(frames[k], seq.poses[k])
            assert b.feed_device(dev[k].data_ptr(), seq.w, seq.h, seq.w * 3, seq.poses[k])
        res = c.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, True)
        assert (res == 0).all()
        for m in (a, b, c):
            m.sync()
        ia, ib, ic = a.get_image(), b.get_image(), c.get_image()
        assert np.array_equal(ia[0], ib[0]) and np.array_equal(ia[0], ic[0])
        assert a.queueSize() == 0
        for m in (a, b, c):
            m.close()


@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("batch", [1, 3, 7, 64])
def test_group_size_does_not_change_results(typ, batch):
    """m2d_feed_batch fuses frames in groups of batch_frames; any grouping must equal sequential feed() calls
    (tile-centric select preserves per-tile frame order, ties included), and must match the oracle bit for bit."""
    import torch
    seq = synth.Sequence(20, 320, 180, seed=12, jitter=True, fpl=5, prepare_frames=3, cross=0.7, along=0.5)
    frames = seq.frames()
    dev = torch.from_numpy(frames).cuda()
    poses = seq.poses.copy()
    poses[9, 3:] = [0.5, 0.5, 0.5, 0.5]  # one rejected frame in the middle of a group
    g = m2d.Map2D.create(typ, thread=False, batch_frames=batch, collect_stats=1)
    o = O.OracleMap2D.create(typ)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    res = g.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, poses, True)
    exp = np.array([0 if o.feed(frames[k], poses[k]) else 1 for k in range(seq.n)])
    assert np.array_equal(res, exp) and res[9] == 1
    g.sync()
    compare_state(g, o, typ)
    sg, so = g.stats(), o.stats()
    for k in ("frames_fed", "frames_fused", "input_px", "region_px", "fresh_px", "win_px"):
        assert sg[k] == so[k], (k, sg[k], so[k])
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("axis,span,count", [(0, 1, 2), (1, 1, 3), (0, 2, 2)])
def test_tile_shards_on_one_gpu_equal_unsharded(typ, axis, span, count):
    """Spatial tile ownership emulated on ONE GPU: `count` shard handles each fuse only the tiles
    they own (multi-band: warped window = owned bbox + one tile ring, clipped to the frame region), then the final
    gather (export -> import into shard 0).  Every tile must be bit-identical to the oracle's unsharded run."""
    import torch
    seq = synth.Sequence(14, 320, 180, seed=21, jitter=True, fpl=4, prepare_frames=3, cross=0.9, along=0.6)
    frames = seq.frames()
    dev = torch.from_numpy(frames).cuda()
    shards = [m2d.Map2D.create(typ, thread=False, shard_rank=r, shard_count=count, shard_axis=axis, shard_span=span, batch_frames=5)
              for r in range(count)]
    o = O.OracleMap2D.create(typ)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(seq.n):
        assert o.feed(frames[k], seq.poses[k])
    for m in shards:
        assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)
        res = m.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, True)
        assert (res == 0).all()
        m.sync()
    counts = [m.tile_count() for m in shards]
    assert sum(c > 0 for c in counts) >= 2, counts
    tb = shards[0].tile_bytes()
    for m in shards[1:]:
        n = m.tile_count()
        buf = torch.empty(max(n, 1) * tb, dtype=torch.uint8, device="cuda")
        xy = m.export_tiles(buf.data_ptr(), n, True)
        assert len(xy) == n
        assert shards[0].import_tiles(xy, buf.data_ptr(), True)
    assert shards[0].tile_count() == sum(counts), "ownership must be disjoint and complete"
    compare_state(shards[0], o, typ)
    for m in shards:
        m.close()


@pytest.mark.parametrize("sparse,cull", [("1", "1"), ("1", "0"), ("0", "1")])
def test_weights_first_and_dense_pipelines_are_bit_exact(monkeypatch, sparse, cull):
    """Multi-band has two pipelines: weights-first (default; competitive cells from closed-form weight bounds, winners
    decided from the weight pyramids, image warp/pyrDown only in the cells a winner's Laplacian needs, DESIGN.md §3;
    M2D_WCULL=0 switches the bound-based culling off) and dense (M2D_SPARSE=0).  Same bits as the oracle from all:
    jittered and noisy frames, several groups, 1/3/5 bands, a sharded window, stats, and the dense fallback the
    weights-first pipeline takes for 8 bands."""
    import torch
    monkeypatch.setenv("M2D_SPARSE", sparse)
    monkeypatch.setenv("M2D_WCULL", cull)
    seq = synth.Sequence(14, 320, 180, seed=29, jitter=True, noise=True, fpl=4, prepare_frames=4)
    dev = torch.from_numpy(seq.frames()).cuda()
    for kw in ({}, {"band_number": 3}, {"band_number": 1}, {"band_number": 8}, {"collect_stats": 1},
               {"shard_rank": 1, "shard_count": 2, "shard_axis": 0, "shard_span": 1}):
        g = m2d.Map2D.create(3, thread=False, batch_frames=5, **kw)
        o = O.OracleMap2D.create(3, **kw)
        assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        res = g.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, True)
        for k in range(seq.n):
            assert (res[k] == 0) == o.feed(seq.frame(k), seq.poses[k])
        g.sync()
        compare_state(g, o, 3)
        if kw.get("collect_stats"):
            assert g.stats()["win_px"] == o.stats()["win_px"]
        g.close()
    # a realistic overlap pattern at full tile scale: 720p serpentine prefix, two groups
    seq = synth.Sequence(24, 1280, 720, seed=2, fpl=6, prepare_frames=6)
    dev = torch.from_numpy(seq.frames()).cuda()
    g = m2d.Map2D.create(3, thread=False, batch_frames=16)
    o = O.OracleMap2D.create(3)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    g.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, True)
    for k in range(seq.n):
        assert o.feed(seq.frame(k), seq.poses[k])
    g.sync()
    compare_state(g, o, 3)
    g.close()


@pytest.mark.parametrize("bands", [7, 8])
def test_sharded_window_ring_follows_the_level_count(bands):
    """The pyramid window of a shard = owned tiles + a ring that must cover the level-0 support of the deepest Laplacian
    tap: 3 * 2^(L-1) - 2 px = 94 for the default 6 levels (1 tile), 382 for 8 levels (2 tiles), 766 for 9 levels (3 tiles).
    Frames wide enough (5 x 4 tiles) for a one-tile ring to be too small: every shard's tiles must still equal the oracle's
    unsharded run."""
    import torch
    seq = synth.Sequence(6, 1100, 800, seed=33, jitter=True, fpl=3, prepare_frames=3, cross=0.5, along=0.4)
    frames = seq.frames()
    dev = torch.from_numpy(frames).cuda()
    o = O.OracleMap2D.create(3, band_number=bands)
    O.set_threads(os.cpu_count() or 1)
    try:
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(seq.n):
            assert o.feed(frames[k], seq.poses[k])
    finally:
        O.set_threads(1)
    count = 3
    shards = [m2d.Map2D.create(3, thread=False, band_number=bands, shard_rank=r, shard_count=count, shard_axis=0, shard_span=1) for r in range(count)]
    for m in shards:
        assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)
        assert (m.feed_batch(dev.data_ptr(), seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, True) == 0).all()
        m.sync()
    tb = shards[0].tile_bytes()
    for m in shards[1:]:
        n = m.tile_count()
        buf = torch.empty(max(n, 1) * tb, dtype=torch.uint8, device="cuda")
        assert shards[0].import_tiles(m.export_tiles(buf.data_ptr(), n, True), buf.data_ptr(), True)
    compare_state(shards[0], o, 3)
    for m in shards:
        m.close()


def raw_state(m):
    """Every tile this handle holds as (sorted absolute coordinates, [n, state_bytes] uint8 CUDA tensor of reference state)."""
    import torch
    n, tb, sb = m.tile_count(), m.tile_bytes(), m.tile_state_bytes()
    buf = torch.empty(max(n, 1) * tb, dtype=torch.uint8, device="cuda")
    xy = m.export_tiles(buf.data_ptr(), n, True)
    order = np.lexsort((xy[:, 0], xy[:, 1]))
    st = buf.view(max(n, 1), tb)[:n, :sb][torch.from_numpy(order).cuda()]
    return xy[order], st


@pytest.mark.parametrize("typ,n,seed", [(1, 100, 1), (3, 500, 2)])
def test_benchmarked_configs_are_bit_exact(typ, n, seed):
    """EXACTLY what bench.py's step_device runs -- BASELINE configs[0] (cfg1: 100 x 1280x720, seed 1, weighted) and
    configs[1] (cfg2: 500 x 1280x720, seed 2, multi-band), frames resident in HBM, ONE m2d_feed_batch call with the
    library's default group size, collect_stats off (best-first order + alpha-bound culling for weighted; competitive-cell
    culling + weights-first for multi-band) -- against the oracle fed frame by frame: every tile of every level and the
    collapsed mosaic, byte for byte.  (Map2DCPU.cpp:324-329, MultiBandMap2DCPU.cpp:539-547)"""
    import torch
    seq = synth.Sequence(n, 1280, 720, seed=seed)
    frames = seq.frames()
    dev = torch.from_numpy(frames).cuda()
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    O.set_threads(os.cpu_count() or 1)
    try:
        assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for rep in range(2):   # the bench resets and feeds again: the second pass runs on recycled pool tiles
            g.reset()
            res = g.feed_batch(dev.data_ptr(), n, 1280 * 720 * 3, 1280, 720, 1280 * 3, seq.poses, True)
        exp = np.array([0 if o.feed(frames[k], seq.poses[k]) else 1 for k in range(n)])
        assert np.array_equal(res, exp)
        g.sync()
        assert compare_state(g, o, typ) > 150
    finally:
        O.set_threads(1)
    g.close()


SWEEP = [  # (w, h, frames, tilt rad, scale, weight_type): >= 1e9 input px in total
    (1280, 720, 300, 0.0, 1.0, 0), (1280, 720, 160, 0.45, 1.0, 1), (1920, 1080, 120, 0.3, 0.5, 0), (1920, 1080, 60, 0.1, 2.0, 0),
    (4000, 3000, 24, 0.2, 1.0, 0), (4000, 3000, 12, 0.48, 0.7, 1)]


@pytest.mark.parametrize("typ", [1, 3])
def test_culling_never_changes_results(monkeypatch, typ):
    """The shortcuts the benchmarked numbers rest on, against the same library with the shortcuts off, over > 1e9 input px
    of randomly tilted (up to 28 deg), yawed, jittered frames at 720p / 1080p / 12 MP and map scales 0.5 - 2:
      weighted   best-first order + alpha upper bounds (collect_stats=0)  ==  sequential, unculled (collect_stats=1)
      multi-band competitive-cell culling (bounds.h) + weights-first     ==  M2D_WCULL=0  ==  dense pipeline (M2D_SPARSE=0)
    Raw tile state (every level, weights included) and the grid must be identical."""
    import torch
    total = 0
    for (w, h, n, tilt, scale, wt) in SWEEP:
        seq = synth.Sequence(n, w, h, seed=31 + n, jitter=True)
        rng = np.random.default_rng(n)
        poses = seq.poses.copy()
        for k in range(n):
            q = synth._qmul(synth._qmul(synth._qaxis((0, 0, 1), rng.uniform(-3.1, 3.1)),
                                        synth._qmul(synth._qaxis((0, 1, 0), rng.uniform(-tilt, tilt)), synth._qaxis((1, 0, 0), rng.uniform(-tilt, tilt)))),
                            np.array([1.0, 0, 0, 0]))
            poses[k, 3:] = q / np.linalg.norm(q)
        tex = torch.from_numpy(seq.texture).cuda()
        dev = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
        for k in range(n):   # the same crops Sequence.frame() makes, gathered on the GPU
            r0, c0 = seq.frame_origin(k)
            rows = torch.from_numpy((r0 - np.arange(h)) % tex.shape[0]).cuda()
            cols = torch.from_numpy((c0 + np.arange(w)) % tex.shape[1]).cuda()
            dev[k] = tex[rows[:, None], cols[None, :]]
        if typ == 1:
            variants = [({"collect_stats": 0}, {}), ({"collect_stats": 1}, {})]
        else:
            variants = [({}, {"M2D_WCULL": "1", "M2D_SPARSE": "1"}), ({}, {"M2D_WCULL": "0", "M2D_SPARSE": "1"}), ({}, {"M2D_WCULL": "1", "M2D_SPARSE": "0"})]
        ref = None
        for kw, env in variants:
            for k_, v_ in env.items():
                monkeypatch.setenv(k_, v_)
            m = m2d.Map2D.create(typ, thread=False, scale=scale, weight_type=wt, **kw)
            assert m.prepare(seq.plane, seq.camera, poses[:20])
            res = m.feed_batch(dev.data_ptr(), n, w * h * 3, w, h, w * 3, poses, True)
            m.sync()
            xy, st = raw_state(m)
            gr = m.grid()
            if ref is None:
                ref = (res, xy, st, gr)
                assert (res == 0).sum() > n // 2
            else:
                assert np.array_equal(res, ref[0]) and np.array_equal(xy, ref[1]) and (gr["w"], gr["h"]) == (ref[3]["w"], ref[3]["h"])
                same = torch.equal(st, ref[2])
                assert same, "%dx%d tilt %.2f scale %.1f: %d tile bytes differ with %s %s" % (
                    w, h, tilt, scale, int((st != ref[2]).sum()), kw, env)
            m.close()
            del m
        total += int((ref[0] == 0).sum()) * w * h
        del dev, ref
        torch.cuda.empty_cache()
    assert total >= 1_000_000_000, total


def test_ties_keep_reference_order():
    """Exact weight ties: the same frame fed twice.  Weighted keeps the first ('<'), multi-band takes the last
    ('>=') -- observable through the win counters; state must stay identical to the oracle either way."""
    seq = synth.Sequence(3, 320, 180, seed=13, fpl=3, prepare_frames=3)
    for typ in (1, 3):
        for batch in (1, 8):
            g = m2d.Map2D.create(typ, thread=False, batch_frames=batch, collect_stats=1)
            o = O.OracleMap2D.create(typ)
            assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
            frames = np.stack([seq.frame(0), seq.frame(0), seq.frame(1), seq.frame(1), seq.frame(0)])
            poses = np.stack([seq.poses[0], seq.poses[0], seq.poses[1], seq.poses[1], seq.poses[0]])
            import torch
            dev = torch.from_numpy(frames).cuda()
            g.feed_batch(dev.data_ptr(), 5, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, poses, True)
            for k in range(5):
                assert o.feed(frames[k], poses[k])
            g.sync()
            compare_state(g, o, typ)
            assert g.stats()["win_px"] == o.stats()["win_px"]
            g.close()


def test_idempotence_and_order_property():
    """Size-independent properties at BASELINE frame size: feeding the same frame twice changes nothing in
    weighted mode (strict '<'), and the mosaic alpha equals the per-pixel max of the individual frames' alphas."""
    seq = synth.Sequence(6, 1280, 720, seed=6, jitter=True, fpl=3, prepare_frames=6)
    g = m2d.Map2D.create(1, thread=False)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(seq.n):
        assert g.feed(seq.frame(k), seq.poses[k])
    first, org = g.get_image()
    for k in range(seq.n):
        assert g.feed(seq.frame(k), seq.poses[k])
    again, org2 = g.get_image()
    assert org == org2 and np.array_equal(first, again)
    singles = []
    for k in range(seq.n):
        s = m2d.Map2D.create(1, thread=False)
        assert s.prepare(seq.plane, seq.camera, seq.prepare_poses)
        assert s.feed(seq.frame(k), seq.poses[k])
        img, o1 = s.get_image()
        full = np.zeros(first.shape[:2], np.uint8)
        y0, x0 = (o1[1] - org[1]) * 256, (o1[0] - org[0]) * 256
        full[y0:y0 + img.shape[0], x0:x0 + img.shape[1]] = img[..., 3]
        singles.append(full)
        s.close()
    assert np.array_equal(first[..., 3], np.max(singles, axis=0))
    g.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_checkpoint_resume_is_transparent(tmp_path, typ):
    """Feed half the survey, m2d_save_state, restore into a NEW handle, feed the rest: every tile must equal the
    uninterrupted oracle run (the grid after spreadMap, the tile origin and all pyramid levels are part of the state)."""
    seq = synth.Sequence(14, 320, 180, seed=29, jitter=True, fpl=4, prepare_frames=2, cross=1.2, along=0.8)
    o = O.OracleMap2D.create(typ)
    a = m2d.Map2D.create(typ, thread=False)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses) and a.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(7):
        assert o.feed(seq.frame(k), seq.poses[k]) and a.feed(seq.frame(k), seq.poses[k])
    path = str(tmp_path / "state.m2d")
    assert a.save_state(path)
    a.close()
    b = m2d.Map2D.create(typ, thread=False)
    assert b.load_state(path)                       # no prepare(): the grid comes from the file
    assert b.grid()["w"] == o.grid()["w"] and np.array_equal(b.grid()["min"], o.grid()["min"])
    for k in range(7, seq.n):
        assert o.feed(seq.frame(k), seq.poses[k]) and b.feed(seq.frame(k), seq.poses[k])
    b.sync()
    compare_state(b, o, typ)
    other = m2d.Map2D.create(4 - typ, thread=False)  # wrong type must be refused loudly
    with pytest.raises(RuntimeError):
        other.load_state(path)
    other.close()
    b.close()


@pytest.mark.parametrize("typ", [1, 3])
def test_display_tiles_match_reference_blend(typ):
    """Map2D::draw without GL: changed-tile polling (Ele::Ischanged) and per-tile textures.  Multi-band tiles with all 8
    neighbours present are collapsed with the borrowed borders of Ele::blend (MultiBandMap2DCPU.cpp:77-146), the others
    alone; both must equal the oracle's restatement byte for byte."""
    seq = synth.Sequence(5, 1024, 768, seed=7, jitter=True, fpl=3, prepare_frames=5)
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert g.poll_changed() == []
    for k in range(3):
        assert g.feed(seq.frame(k), seq.poses[k]) and o.feed(seq.frame(k), seq.poses[k])
    first = g.poll_changed()
    assert len(first) == g.tile_count() and g.poll_changed() == []   # flags are cleared by the poll
    for k in range(3, seq.n):
        assert g.feed(seq.frame(k), seq.poses[k]) and o.feed(seq.frame(k), seq.poses[k])
    x0, y0, x1, y1 = g.last_rect()
    again = g.poll_changed()
    assert (x0, y0) in again and len(again) >= (x1 - x0) * (y1 - y0)
    gr = o.grid()
    hq_differs = 0
    for ty in range(gr["h"]):
        for tx in range(gr["w"]):
            for hq in (True, False):
                a, b = o.get_tile_image(tx, ty, hq), g.get_tile_image(tx, ty, hq)
                assert (a is None) == (b is None)
                if a is not None:
                    assert np.array_equal(a, b), "tile (%d,%d) hq=%s: %d bytes differ" % (tx, ty, hq, int((a != b).sum()))
            a, b = g.get_tile_image(tx, ty, True), g.get_tile_image(tx, ty, False)
            if a is not None and not np.array_equal(a, b):
                hq_differs += 1
    if typ == 3:
        assert hq_differs >= 4, "the survey must contain interior tiles (all 8 neighbours present)"
    g.close()


def test_save_png_roundtrip(tmp_path):
    cv2 = pytest.importorskip("cv2")
    seq = synth.Sequence(4, 320, 180, seed=1, fpl=2, prepare_frames=4)
    for typ in (1, 3):
        g = m2d.Map2D.create(typ, thread=False)
        assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(seq.n):
            assert g.feed(seq.frame(k), seq.poses[k])
        path = str(tmp_path / ("m%d.png" % typ))
        assert g.save(path)
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        assert np.array_equal(img, g.get_image()[0])
        g.close()
