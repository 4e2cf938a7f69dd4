"""The sharded driver on real GPUs (NCCL, one process per GPU): needs >= 2 CUDA devices, skipped otherwise (the round-end
GPU box has one; run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`).  Same checks as the
gloo test, but the per-rank map is the CUDA library and the result on rank 0 is compared with the CPU oracle."""
import pytest
import torch
import torch.multiprocessing as mp

from tests.test_sharded_gloo import _free_port, _worker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("delivery", [False, True, "owned", "peer"])
def test_two_gpu_sharded_equals_oracle(typ, delivery):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, typ, delivery, q, "nccl")) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    root = [o for o in outs if o["rank"] == 0][0]
    other = [o for o in outs if o["rank"] == 1][0]
    assert root["res"] == other["res"] == root["exp"] and root["res"][6] == 1
    assert root["same_grid"] and min(root["counts"]) > 0
    assert root["received"] == root["counts"][1] and root["ntiles"] == sum(root["counts"])
    assert root["bad"] == 0 and root["image_equal"]


@pytest.mark.parametrize("typ", [1, 3])
def test_multi_device_handle_equals_oracle(typ):
    """m2d_create_multi: ONE process, two GPUs behind one handle (block-cyclic strips of one tile).  Host frames through
    feed(), then device-resident frames (on GPU 0, read in place over NVLink by GPU 1) through one feed_batch; grid, every
    tile of every level (fetched from the owning device) and the saved mosaic must equal the oracle's, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    import pi_slam_fusion_b200.map2d as m2d
    import pi_slam_fusion_b200.synth as synth
    from oracle import oracle as O
    from tests.test_parity_gpu import compare_state
    seq = synth.Sequence(24, 320, 180, seed=21, jitter=True, fpl=4, prepare_frames=3, cross=0.9, along=0.6)
    frames = seq.frames()
    g = m2d.Map2D.create(typ, thread=False, devices=[0, 1], shard_axis=0, shard_span=1)
    o = O.OracleMap2D.create(typ)
    assert g.prepare(seq.plane, seq.camera, seq.prepare_poses) and o.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for k in range(8):
        assert g.feed(frames[k], seq.poses[k]) == o.feed(frames[k], seq.poses[k])
        assert g.last_rect() == o.last_rect()
    dev = torch.from_numpy(frames[8:]).cuda(0)
    res = g.feed_batch(dev.data_ptr(), seq.n - 8, seq.w * seq.h * 3, seq.w, seq.h, seq.w * 3, seq.poses[8:], True)
    exp = [0 if o.feed(frames[k], seq.poses[k]) else 1 for k in range(8, seq.n)]
    assert res.tolist() == exp
    g.sync()
    assert g.queueSize() == 0 and g.tile_count() == o.tile_count()
    compare_state(g, o, typ)
    n_before = g.tile_count()
    compare_state(g, o, typ)          # the save gave the gathered copies back: a second save sees the same map
    assert g.tile_count() == n_before
    g.close()
