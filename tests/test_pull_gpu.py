"""Pinned host frames are not staged whole (csrc/map2d.cu feed_frames, kernels_wf.cu mbs_mark / mbs_pull): the weights-first
multi-band pipeline pulls only the 256-byte chunks of every frame that its winners' cells sample, the weighted kernel samples
the host frames in place over PCIe.  Results must not change by a bit.  M2D_PULL_POISON=1 fills the staging slots with 0xA5
before every pull, so a chunk the marking missed cannot go unnoticed (the frames are i.i.d. noise)."""
import numpy as np
import pytest

import pi_slam_fusion_b200.map2d as m2d
import pi_slam_fusion_b200.synth as synth
from oracle import oracle as O
from tests.test_parity_gpu import compare_state

pytestmark = pytest.mark.gpu


def pinned_frames(seq):
    host, ptr = m2d.pinned_empty((seq.n, seq.h, seq.w, 3))
    for k in range(seq.n):
        host[k] = seq.frame(k)
    return host, ptr


def feed_pinned(typ, seq, host, ptr, **cfg):
    g = m2d.Map2D.create(typ, thread=False, **cfg)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    res = g.feed_batch(ptr, seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, False)
    g.sync()
    return g, res


def oracle_of(typ, seq, host, **cfg):
    O.set_threads(8)
    o = O.OracleMap2D.create(typ, **cfg)
    assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    acc = [o.feed(host[k], seq.poses[k]) for k in range(seq.n)]
    O.set_threads(1)
    return o, acc


@pytest.mark.parametrize("w,h,scale,tilt", [(640, 360, 1.0, False), (640, 360, 0.6, True), (640, 360, 1.9, True), (328, 180, 1.0, True),
                                            (321, 181, 1.0, False)])
def test_pull_mode_is_bit_exact(monkeypatch, w, h, scale, tilt):
    """(328 x 180: frame bytes not a multiple of 256 -> a partial last chunk; 321 x 181: not a multiple of 16 -> falls back to
    the staging copies.)"""
    monkeypatch.setenv("M2D_PULL_POISON", "1")
    seq = synth.Sequence(24, w, h, seed=31, jitter=True, noise=True, fpl=6, prepare_frames=6)
    if tilt:   # up to ~20 degrees of roll / pitch on every third frame: genuinely projective homographies
        rng = np.random.default_rng(7)
        for k in range(0, seq.n, 3):
            q = synth._qmul(synth._qaxis((1, 0, 0), np.radians(rng.uniform(-20, 20))), synth._qaxis((0, 1, 0), np.radians(rng.uniform(-20, 20))))
            seq.poses[k, 3:] = synth._qmul(q, seq.poses[k, 3:])
    host, ptr = pinned_frames(seq)
    try:
        g, res = feed_pinned(3, seq, host, ptr, scale=scale)
        o, acc = oracle_of(3, seq, host, scale=scale)
        assert [r == 0 for r in res] == acc
        compare_state(g, o, 3)
        pulled = g.launch_count()
        gi, go = g.get_image()
        oi, oo = o.get_image()
        assert go == oo and np.array_equal(gi, oi)
        g.close()
        monkeypatch.setenv("M2D_ZEROCOPY", "0")   # A/B: whole-frame staging copies
        g2, res2 = feed_pinned(3, seq, host, ptr, scale=scale)
        assert np.array_equal(res, res2)
        compare_state(g2, o, 3)
        staged = g2.launch_count()
        g2.close()
        # one group of 24 frames: pull mode adds the mark and the pull launch; a frame size that is not a multiple of 16 bytes
        # keeps the staging copies
        assert pulled == staged + (2 if (w * h * 3) % 16 == 0 else 0), (pulled, staged)
    finally:
        m2d.free_pinned(ptr)


def test_weighted_samples_pinned_frames_in_place():
    seq = synth.Sequence(20, 640, 360, seed=12, jitter=True, noise=True, fpl=5, prepare_frames=5)
    host, ptr = pinned_frames(seq)
    try:
        g, res = feed_pinned(1, seq, host, ptr)
        o, acc = oracle_of(1, seq, host)
        assert [r == 0 for r in res] == acc
        compare_state(g, o, 1)
        g.close()
    finally:
        m2d.free_pinned(ptr)


def test_pageable_frames_and_small_batches_are_staged():
    """numpy memory is pageable and a batch of < 8 frames stays on the staging path: same results."""
    seq = synth.Sequence(12, 640, 360, seed=3, jitter=True, fpl=4, prepare_frames=4)
    frames = seq.frames()
    o, acc = oracle_of(3, seq, frames)
    g = m2d.Map2D.create(3, thread=False)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
    res = g.feed_batch(frames.ctypes.data, seq.n, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses, False)
    g.sync()
    assert [r == 0 for r in res] == acc
    compare_state(g, o, 3)
    g.close()
    host, ptr = pinned_frames(seq)
    try:
        g = m2d.Map2D.create(3, thread=False)
        assert g.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for i in range(0, seq.n, 6):   # pinned, but 6 frames per call
            g.feed_batch(ptr + i * seq.w * seq.h * 3, 6, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses[i:i + 6], False)
        g.sync()
        compare_state(g, o, 3)
        g.close()
    finally:
        m2d.free_pinned(ptr)


@pytest.mark.parametrize("typ", [1, 3])
def test_sparse_surveys_keep_the_staging_copies(monkeypatch, typ):
    """Frames that barely overlap (5 % side lap: < 2.5 frames per touched tile) are read about once either way, and the copy
    engine moves whole frames faster than the SMs can pull them: such groups stay on the staging path.  Same results."""
    seq = synth.Sequence(12, 640, 360, seed=3, jitter=True, noise=True, along=0.95, cross=0.95, fpl=4, prepare_frames=4)
    host, ptr = pinned_frames(seq)
    try:
        g, res = feed_pinned(typ, seq, host, ptr)
        o, acc = oracle_of(typ, seq, host)
        assert [r == 0 for r in res] == acc
        compare_state(g, o, typ)
        a = g.launch_count()
        g.close()
        monkeypatch.setenv("M2D_ZEROCOPY", "0")
        g2, _ = feed_pinned(typ, seq, host, ptr)
        compare_state(g2, o, typ)
        assert a == g2.launch_count()        # no mark / pull launches were added
        g2.close()
    finally:
        m2d.free_pinned(ptr)
