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


def _worker(rank, world, port, typ, distributed_input, q, backend="gloo"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    cuda = backend == "nccl"
    if cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        import pi_slam_fusion_b200.synth as synth
        from pi_slam_fusion_b200.sharded import ShardedMap2D
        if cuda:   # the product: the CUDA library on this rank's GPU (the oracle only checks the result on rank 0)
            import pi_slam_fusion_b200.map2d as m2d
            factory, device, tdev = (lambda t, **kw: m2d.Map2D.create(t, thread=False, **kw)), rank, torch.device("cuda", rank)
        else:
            factory, device, tdev = (lambda t, **kw: O.OracleMap2D.create(t, **kw)), None, torch.device("cpu")
        from pi_slam_fusion_b200.sharded import DeliveryPlan, even_split
        if distributed_input in ("owned", "peer"):   # a longer strip survey, so that some frames are needed by one rank only
            seq = synth.Sequence(30, 320, 180, seed=23, jitter=True, fpl=3, prepare_frames=3, cross=0.9, along=0.6)
        else:
            seq = synth.Sequence(14, 320, 180, seed=21, jitter=True, fpl=4, prepare_frames=3, cross=0.9, along=0.6)
        frames = torch.from_numpy(seq.frames()).to(tdev) if rank == 0 else None
        sm = ShardedMap2D(factory, typ, rank, world, device=device, shard_axis=0, shard_span=1)
        assert sm.prepare(seq.plane, seq.camera, seq.prepare_poses)
        poses = seq.poses.copy()
        poses[6, 3:] = [0.5, 0.5, 0.5, 0.5]  # rejected on every rank alike
        extra = {}
        if distributed_input in ("owned", "peer"):
            rects, axis, span, origin = sm.align_strips(poses)
            plan = DeliveryPlan(rects, axis, span, world, even_split(seq.n, world), origin)
            lo, hi = plan.resident[rank]
            if distributed_input == "peer":   # CUDA only: halo frames are sampled in place from the neighbour's HBM (CUDA IPC)
                mine, ptrs = sm.share_frames(plan, seq.w, seq.h)
                mine.copy_(torch.from_numpy(seq.frames(range(lo, hi))).to(tdev))
                torch.cuda.synchronize()
                dist.barrier()
                res = sm.feed_all_peer(plan, ptrs, poses, seq.w, seq.h)
            else:
                buf, mine = sm.alloc_owned_buffer(plan, seq.w, seq.h)
                buf.fill_(7)  # halo slots start as garbage: they must be overwritten by the exchange
                mine.copy_(torch.from_numpy(seq.frames(range(lo, hi))))
                res = sm.feed_all_owned(plan, buf, poses, seq.w, seq.h)
            extra = {"hull": plan.hull, "moved": plan.frames_moved(), "n": seq.n}
            # sharded save BEFORE any gather: every rank collapses its own strip (+ one halo tile row from its neighbours)
            before = sm.map.tile_count()
            mosaic, gbox = sm.save_sharded(axis, span, origin, gather=True)
            assert sm.map.tile_count() == before, "halo tiles must be dropped again"
            strip, srect = sm.save_sharded(axis, span, origin)      # and without the final gather: just my strip
            extra["strip"] = None if strip is None else (tuple(strip.shape), tuple(srect))
            if rank == 0:
                extra["sharded_mosaic"] = mosaic.cpu().numpy()
                extra["gbox"] = tuple(gbox)
        elif distributed_input:
            ids = ShardedMap2D.local_frame_ids(seq.n, rank, world, block=3)
            local = torch.from_numpy(seq.frames(ids)).to(tdev)
            res = sm.feed_all_distributed(local, poses, seq.w, seq.h, block=3)
        else:
            res = sm.feed_all(frames, poses, seq.w, seq.h, chunk=5)
        owned = sm.map.tile_count()
        counts = [None] * world
        dist.all_gather_object(counts, owned)
        got = sm.gather_to_root()
        out = {"rank": rank, "res": res.tolist(), "counts": counts, "received": got}
        out.update(extra)
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
            if "sharded_mosaic" in out:   # the strips of the sharded save tile the reference's mosaic byte for byte
                sh = out.pop("sharded_mosaic")
                out["sharded_save_equal"] = bool(sh.shape == ia[0].shape and np.array_equal(sh, ia[0]))
                out["sharded_save_bbox_ok"] = bool(ref.tile_bbox() == out["gbox"])
        if distributed_input == "peer":
            sm.unshare_frames()
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("typ,distributed_input,world", [(1, False, 2), (3, False, 2), (1, True, 2), (3, True, 2),
                                                         (1, "owned", 2), (3, "owned", 2), (3, "owned", 3)])
def test_two_rank_sharded_equals_unsharded(typ, distributed_input, world):
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
    for other in outs:
        assert other["res"] == root["exp"]
    assert root["res"][6] == 1
    assert root["same_grid"]
    assert min(root["counts"]) > 0, "both shards must own tiles in this layout"
    assert root["received"] == sum(root["counts"][1:])
    assert root["ntiles"] == sum(root["counts"]), "ownership must be disjoint and complete"
    assert root["bad"] == 0 and root["image_equal"]
    if distributed_input == "owned":
        # the plan really kept pixels away from ranks that do not need them, and really moved the halo frames
        assert any(tuple(h) != (0, root["n"]) for h in root["hull"]) and 0 < root["moved"] < root["n"]
        assert root["sharded_save_equal"] and root["sharded_save_bbox_ok"]
        strips = [o["strip"] for o in sorted(outs, key=lambda o: o["rank"]) if o["strip"] is not None]
        assert len(strips) >= 2, "at least two ranks must hold a strip of the mosaic"


def test_delivery_plan_properties():
    import sys
    sys.path.insert(0, ROOT)
    from pi_slam_fusion_b200.sharded import DeliveryPlan, even_split, strip_owner
    rng = np.random.default_rng(5)
    n, world, span, origin = 60, 4, 6, -7
    x0 = np.sort(rng.integers(-7, 14, n))
    rects = np.stack([x0, rng.integers(0, 3, n), x0 + rng.integers(1, 5, n), rng.integers(4, 6, n)], 1)
    rects[11] = -1
    plan = DeliveryPlan(rects, 0, span, world, even_split(n, world), origin, margin=1)
    exact = DeliveryPlan(rects, 0, span, world, even_split(n, world), origin)
    assert all(e[0] >= h[0] and e[1] <= h[1] for e, h in zip(exact.hull, plan.hull))
    for r in range(world):
        a, b = plan.hull[r]
        for k in range(n):
            owners = {strip_owner(t, span, world, origin) for t in range(rects[k, 0], rects[k, 2])} if k != 11 else set()
            if r in owners:
                assert a <= k < b, "a frame that touches an owned tile must be inside the rank's hull"
        lo, hi = plan.buffer[r]
        assert lo <= min(a, plan.resident[r][0]) and hi >= max(b, plan.resident[r][1])
        got = sorted([(t[2], t[3]) for t in plan.transfers if t[1] == r] + [plan.resident[r]])
        cover = [k for (u, v) in got for k in range(max(u, a), min(v, b))]
        assert cover == list(range(a, b)), "hull = own frames + received ranges, without overlap"
    assert strip_owner(origin, span, world, origin) == 0 and strip_owner(origin - 1, span, world, origin) == world - 1


def test_oracle_plan_rects_equal_sequential_feed_rects():
    """The CPU stand-in's dry run (orc_plan_rects) against what feeding the frames one by one really does."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    import pi_slam_fusion_b200.synth as synth
    seq = synth.Sequence(26, 320, 180, seed=31, jitter=True, fpl=3, prepare_frames=3, cross=0.9, along=0.6)
    poses = seq.poses[::-1].copy()   # fly the survey backwards: the grid then also grows towards negative coordinates
    poses[4, 3:] = [0.5, 0.5, 0.5, 0.5]
    o = O.OracleMap2D.create(1)
    assert o.prepare(seq.plane, seq.camera, poses[:3])
    plan = o.plan_rects(poses)
    g0 = o.grid()
    ref = O.OracleMap2D.create(1)
    assert ref.prepare(seq.plane, seq.camera, poses[:3])
    org = np.zeros(2)
    m0 = ref.grid()["min"][:2].copy()
    es = 256 * ref.grid()["length_pixel"]
    grew = False
    for k in range(seq.n):
        ok = ref.feed(seq.frame(seq.n - 1 - k), poses[k])
        if not ok:
            assert (plan[k] == -1).all()
            continue
        g = ref.grid()
        org = np.round((g["min"][:2] - m0) / es).astype(int)   # absolute coordinate of slot (0,0) after the growth
        grew |= bool(org.any())
        r = np.array(ref.last_rect())
        assert (plan[k] == r + [org[0], org[1], org[0], org[1]]).all(), k
    assert grew, "the sequence must exercise spreadMap towards negative coordinates"
    g1 = o.grid()
    assert (g0["w"], g0["h"]) == (g1["w"], g1["h"]) and o.tile_count() == 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_bench_scale_delivery_plan(world):
    """The plan `bench.py --gpus N` builds (N x the 500-frame 720p survey, strips of tiles, frames resident on the
    rank of their own flight lines), computed with the CPU stand-in: every rank's hull is a modest superset of its own
    frames, transfers only connect neighbouring ranks, and every needed frame is either resident or received."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    import pi_slam_fusion_b200.synth as synth
    from pi_slam_fusion_b200.sharded import DeliveryPlan, ShardedMap2D, even_split
    W, H, per_gpu = 1280, 720, 500
    n = per_gpu * world
    poses = synth.serpentine_poses(n, W, H, seed=2, fpl=synth.frames_per_line(per_gpu, W, H))
    cam = synth.camera(W, H)
    plans = []
    for rank in (0, world - 1):
        sm = ShardedMap2D(lambda t, **kw: O.OracleMap2D.create(t, **kw), 3, rank, world, device=None, shard_axis=0, shard_span=4)
        assert sm.prepare(synth.IDENTITY_POSE, cam, poses[:20])
        rects, axis, span, origin = sm.align_strips(poses)
        plans.append(DeliveryPlan(rects, axis, span, world, even_split(n, world), origin))
        # pose-only feeding of everything outside the hull must be accepted by the stand-in (no owned tile under it)
        a, b = plans[-1].hull[rank]
        assert (sm.map.feed_poses(poses[:a]) == 0).all() and (sm.map.feed_poses(poses[b:]) == 0).all()
    p = plans[0]
    assert p.hull == plans[1].hull and p.transfers == plans[1].transfers, "every rank must derive the same plan"
    assert axis == 0
    for r in range(world):
        a, b = p.hull[r]
        lo, hi = p.resident[r]
        # equal frame counts (residency) and equal tile spans (strips) drift apart by up to ~2 flight lines of 42 at N=8
        assert a <= lo + 100 and b >= hi - 100, "strips and residency must roughly coincide"
        assert b - a <= 1.4 * per_gpu, "hull of rank %d is %d frames" % (r, b - a)
    assert all(abs(s - d) == 1 for (s, d, _, _) in p.transfers), "only neighbouring strips exchange frames"
    assert p.frames_moved() <= 0.35 * n
