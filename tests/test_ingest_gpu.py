"""Ingest seam (reference src/DataTrans.h:54-85, TrackerOpt.cpp:374-383, Map2DCPU.cpp:139-142):
bounded drop-oldest queue + BGRA->BGR + worker thread in front of feed(), through the C-ABI on the GPU."""
import threading

import numpy as np
import pytest

import pi_slam_fusion_b200.map2d as m2d
import pi_slam_fusion_b200.synth as synth
from oracle import oracle as O
from tests.test_parity_gpu import compare_state

pytestmark = pytest.mark.gpu


def bgra(img, seed):
    """The tracker's CV_8UC4 frame: same B,G,R, arbitrary 4th channel, and a padded row pitch."""
    h, w, _ = img.shape
    buf = np.random.default_rng(seed).integers(0, 256, (h, w + 3, 4), dtype=np.uint8)
    buf[:, :w, :3] = img
    return buf[:, :w, :]


@pytest.mark.parametrize("typ", [1, 3])
def test_ingest_drop_oldest_bgra_and_order(typ):
    seq = synth.Sequence(12, 320, 180, seed=33, jitter=True, fpl=4, prepare_frames=4)
    frames = seq.frames()
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert g.ingest_open(capacity=5, start_paused=True)
    for k in range(9):   # queue of 5: frames 0..3 fall off the front (DataTrans::product pops the oldest while full)
        assert g.ingest_push(bgra(frames[k], k) if k % 2 else frames[k], seq.poses[k])
    assert not g.ingest_push(np.zeros((100, 320, 3), np.uint8), seq.poses[0])   # wrong frame size: rejected at the seam
    st = g.ingest_stats()
    assert (st["pushed"], st["dropped"], st["fed"]) == (9, 4, 0) and g.queueSize() == 5
    with pytest.raises(RuntimeError):
        g.ingest_drain()          # paused: draining would never return
    g.ingest_pause(False)
    assert g.ingest_drain()
    st = g.ingest_stats()
    assert (st["fed"], st["fused"]) == (5, 5) and g.queueSize() == 0
    for k in range(4, 9):
        assert o.feed(frames[k], seq.poses[k])
    compare_state(g, o, typ)
    # keep streaming while running; an oblique pose is consumed and rejected by the worker, like renderFrame does
    bad = seq.poses[9].copy()
    bad[3:] = [0.5, 0.5, 0.5, 0.5]
    assert g.ingest_push(frames[9], bad)
    for k in range(9, 12):
        assert g.ingest_push(bgra(frames[k], k), seq.poses[k])
        assert o.feed(frames[k], seq.poses[k])
    assert g.ingest_drain()
    st = g.ingest_stats()
    assert st["pushed"] == 13 and st["fed"] + st["dropped"] == 13 and st["fused"] == st["fed"] - 1
    if st["dropped"] == 4:       # nothing more was lost: bit-identical to the oracle fed the same frames
        compare_state(g, o, typ)
    assert g.ingest_close()
    assert g.feed(frames[0], seq.poses[0]) and o.feed(frames[0], seq.poses[0])   # the handle keeps working without the queue
    g.sync()
    if st["dropped"] == 4:
        compare_state(g, o, typ)
    g.close()


def test_ingest_producer_thread_with_concurrent_readers():
    """A producer thread pushes while the main thread polls queueSize() and reads the mosaic: calls are serialised with
    the worker, nothing is lost with a roomy queue, the final state equals the oracle."""
    seq = synth.Sequence(16, 320, 180, seed=35, jitter=True, fpl=4, prepare_frames=4)
    frames = seq.frames()
    g = m2d.Map2D.create(3, thread=False)
    o = O.OracleMap2D.create(3)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    assert g.ingest_open(capacity=30)

    def producer():
        for k in range(seq.n):
            g.ingest_push(frames[k], seq.poses[k])

    t = threading.Thread(target=producer)
    t.start()
    seen = 0
    while t.is_alive() or g.queueSize() > 0:
        img = g.get_image()
        seen += img is not None
        g.poll_changed()
    t.join()
    assert g.ingest_drain()
    st = g.ingest_stats()
    assert st == {"pushed": 16, "dropped": 0, "fed": 16, "fused": 16}
    for k in range(seq.n):
        assert o.feed(frames[k], seq.poses[k])
    compare_state(g, o, 3)
    g.close()   # destroy closes the queue
