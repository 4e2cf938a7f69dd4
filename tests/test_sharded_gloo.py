"""N>1 host logic on CPU: world_size-2 gloo run of the sharded driver (pi-slam-fusion_b200/sharded.py) with the CPU
oracle injected as the per-rank map (the product injects the CUDA Map2D).  Checks frame broadcast, spatial tile
ownership (disjoint, complete), the one-tile-ring multi-band window (bit-identical tiles) and the final gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, typ, distributed_input, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        import pi_slam_fusion_b200.synth as synth
        from pi_slam_fusion_b200.sharded import ShardedMap2D
        seq = synth.Sequence(14, 320, 180, seed=21, jitter=True, fpl=4, prepare_frames=3, cross=0.9, along=0.6)
        frames = torch.from_numpy(seq.frames()) if rank == 0 else None
        sm = ShardedMap2D(lambda t, **kw: O.OracleMap2D.create(t, **kw), typ, rank, world, device=None, shard_axis=0, shard_span=1)
        assert sm.prepare(seq.plane, seq.camera, seq.prepare_poses)
        poses = seq.poses.copy()
        poses[6, 3:] = [0.5, 0.5, 0.5, 0.5]  # rejected on every rank alike
        if distributed_input:
            ids = ShardedMap2D.local_frame_ids(seq.n, rank, world, block=3)
            local = torch.from_numpy(seq.frames(ids))
            res = sm.feed_all_distributed(local, poses, seq.w, seq.h, block=3)
        else:
            res = sm.feed_all(frames, poses, seq.w, seq.h, chunk=5)
        owned = sm.map.tile_count()
        counts = [None] * world
        dist.all_gather_object(counts, owned)
        got = sm.gather_to_root()
        out = {"rank": rank, "res": res.tolist(), "counts": counts, "received": got}
        if rank == 0:
            ref = O.OracleMap2D.create(typ)
            assert ref.prepare(seq.plane, seq.camera, seq.prepare_poses)
            full = seq.frames()
            exp = [0 if ref.feed(full[k], poses[k]) else 1 for k in range(seq.n)]
            g, gr = sm.map.grid(), ref.grid()
            same_grid = (g["w"], g["h"]) == (gr["w"], gr["h"]) and np.array_equal(g["min"], gr["min"])
            ntiles, bad = 0, 0
            for ty in range(gr["h"]):
                for tx in range(gr["w"]):
                    levels = ref.levels if typ == 3 else 1
                    for l in range(levels):
                        a, b = ref.get_tile(tx, ty, l), sm.map.get_tile(tx, ty, l)
                        if (a is None) != (b is None):
                            bad += 1
                            break
                        if a is None:
                            break
                        ntiles += (l == 0)
                        if typ == 3:
                            bad += int(not (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])))
                        else:
                            bad += int(not np.array_equal(a, b))
            ia, ib = ref.get_image(), sm.map.get_image()
            out.update(exp=exp, same_grid=bool(same_grid), ntiles=ntiles, bad=bad,
                       image_equal=bool(ia[1] == ib[1] and np.array_equal(ia[0], ib[0])))
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("distributed_input", [False, True])
def test_two_rank_sharded_equals_unsharded(typ, distributed_input):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, typ, distributed_input, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    root = [o for o in outs if o["rank"] == 0][0]
    other = [o for o in outs if o["rank"] == 1][0]
    assert root["res"] == other["res"] == root["exp"] and root["res"][6] == 1
    assert root["same_grid"]
    assert min(root["counts"]) > 0, "both shards must own tiles in this layout"
    assert root["received"] == root["counts"][1]
    assert root["ntiles"] == sum(root["counts"]), "ownership must be disjoint and complete"
    assert root["bad"] == 0 and root["image_equal"]
