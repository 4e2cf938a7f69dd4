"""Tile-sharded Map2D over the GPUs of one node: one process per GPU, torch.distributed plumbing.

Ownership is spatial: tile (ax, ay) in ABSOLUTE tile coordinates belongs to rank  floor(a_axis / span) mod world
(m2d_config.shard_*).  The only exchange steps of the path are
  * frame delivery  — rank 0 holds the frames; each chunk is broadcast (NCCL over NVLink), every rank sees every
    pose (so all ranks take identical prepare/spreadMap decisions) and fuses only the tiles it owns; for
    multi-band it warps its owned window plus a one-tile ring, so its tiles are bit-identical to an unsharded run;
  * final tile gather — raw tile state to rank 0 (m2d_export_tiles / m2d_import_tiles), which then collapses/saves.
No collective touches the fusion itself.  The map class is injected (`factory`) so that the host logic can be tested
with world_size-2 gloo on a CPU-only box against a CPU stand-in; the product path always injects map2d.Map2D.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def default_shard(w, h, world, poses=None, camera=None):
    """Contiguous strips along the axis on which the survey is longer, sized so that every rank owns about one strip
    (strip count ~ world): balanced load with the least ring recomputation.  Falls back to one frame width."""
    fw, fh = int(np.ceil(w / 256.0)) + 1, int(np.ceil(h / 256.0)) + 1
    if poses is None or camera is None or len(poses) < 2:
        return {"shard_axis": 0, "shard_span": max(2, fw)}
    poses = np.asarray(poses, np.float64).reshape(-1, 7)
    gsd = float(np.median(np.abs(poses[:, 2]))) / float(camera[2])          # ground size of a map px at scale 1
    ext = [(poses[:, 0].max() - poses[:, 0].min()) / (256.0 * gsd) + fw, (poses[:, 1].max() - poses[:, 1].min()) / (256.0 * gsd) + fh]
    axis = 0 if ext[0] >= ext[1] else 1
    return {"shard_axis": axis, "shard_span": max(2, int(np.ceil(ext[axis] / world)))}


class ShardedMap2D:
    def __init__(self, factory, type_, rank, world, device=None, **cfg):
        self.rank, self.world, self.type = rank, world, type_
        self.cuda = device is not None
        self.device = device
        kw = dict(cfg)
        kw.update(shard_rank=rank, shard_count=world)
        if self.cuda:
            kw["device"] = device  # the library must run on the rank's own GPU (m2d_config.device)
        self.map = factory(type_, **kw)

    # Map2D::prepare — identical on every rank
    def prepare(self, plane, camera, poses):
        return self.map.prepare(plane, camera, poses)

    def feed_all(self, frames, poses, w, h, chunk=128):
        """frames: uint8 tensor [n,h,w,3] on rank 0 (CUDA tensor for NCCL, CPU tensor for gloo); other ranks pass None.
        Chunks are double-buffered: the broadcast of chunk c+1 is in flight while the library fuses chunk c.
        Returns the per-frame status array (identical on all ranks)."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        res = np.zeros(n, np.int32)
        starts = list(range(0, n, chunk))
        bufs = [None, None]

        def start(ci):
            c0 = starts[ci]
            m = min(chunk, n - c0)
            if self.rank == 0:
                buf = frames[c0:c0 + m]
            else:
                if bufs[ci % 2] is None:
                    bufs[ci % 2] = torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=dev)
                buf = bufs[ci % 2][:m]
            work = dist.broadcast(buf, src=0, async_op=True) if self.world > 1 else None
            return buf, work, c0, m

        cur = start(0) if starts else None
        for ci in range(len(starts)):
            buf, work, c0, m = cur
            if work is not None:
                work.wait()
                if self.cuda:
                    torch.cuda.current_stream().synchronize()  # the library consumes buf on its own stream
            if ci + 1 < len(starts):
                if ci >= 1:
                    self.map.sync()  # chunk ci-1 has been consumed: its buffer may be overwritten by chunk ci+1
                nxt = start(ci + 1)
            else:
                nxt = None
            res[c0:c0 + m] = self.map.feed_batch(buf.data_ptr(), m, w * h * 3, w, h, w * 3, poses[c0:c0 + m], self.cuda)
            cur = nxt
        self.map.sync()
        return res

    @staticmethod
    def local_frame_ids(n, rank, world, block=16):
        """Frames resident on `rank` when the inputs are spread block-cyclically over the job's GPUs."""
        return [k for k in range(n) if (k // block) % world == rank]

    def feed_all_distributed(self, local_frames, poses, w, h, block=16):
        """Frames originate on ALL GPUs (block-cyclic, local_frame_ids): rank r holds frames k with (k // block) %
        world == r as a uint8 tensor [n_local,h,w,3].  Each chunk of world*block consecutive frames is assembled on
        every rank with one all_gather (every GPU sends and receives at NVLink speed, instead of one root feeding
        everybody), double-buffered against the fusion of the previous chunk."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        res = np.zeros(n, np.int32)
        per_chunk = block * self.world
        starts = list(range(0, n, per_chunk))
        bufs = [None, None]
        pad = None

        def start(ci):
            c0 = starts[ci]
            if bufs[ci % 2] is None:
                bufs[ci % 2] = torch.empty((per_chunk, h, w, 3), dtype=torch.uint8, device=dev)
            out = bufs[ci % 2]
            lo = ci * block                      # my block of this chunk inside local_frames
            mine = local_frames[lo:lo + block]
            if mine.shape[0] < block:            # ragged tail: pad my contribution
                nonlocal pad
                if pad is None:
                    pad = torch.zeros((block, h, w, 3), dtype=torch.uint8, device=dev)
                pad[:mine.shape[0]] = mine
                mine = pad
            if self.world > 1:
                work = dist.all_gather_into_tensor(out.view(-1), mine.contiguous().view(-1), async_op=True)
            else:
                out[:block] = mine
                work = None
            return out, work, c0, min(per_chunk, n - c0)

        cur = start(0) if starts else None
        for ci in range(len(starts)):
            buf, work, c0, m = cur
            if work is not None:
                work.wait()
                if self.cuda:
                    torch.cuda.current_stream().synchronize()
            if ci + 1 < len(starts):
                if ci >= 1:
                    self.map.sync()
                nxt = start(ci + 1)
            else:
                nxt = None
            res[c0:c0 + m] = self.map.feed_batch(buf.data_ptr(), m, w * h * 3, w, h, w * 3, poses[c0:c0 + m], self.cuda)
            cur = nxt
        self.map.sync()
        return res

    def gather_to_root(self):
        """Raw owned tiles -> rank 0 (which imports them).  Returns the number of tiles received by the root."""
        tb = self.map.tile_bytes()
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        n_local = self.map.tile_count()
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
        if self.world > 1:
            dist.all_gather(counts, mine)
        else:
            counts = [mine]
        counts = [int(c.item()) for c in counts]
        received = 0
        if self.rank != 0:
            buf = torch.empty(max(n_local, 1) * tb, dtype=torch.uint8, device=dev)
            xy = self.map.export_tiles(buf.data_ptr(), n_local, self.cuda)
            xy_t = torch.from_numpy(np.ascontiguousarray(xy.reshape(-1))).to(dev)
            if n_local:
                for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, xy_t, 0), dist.P2POp(dist.isend, buf[:n_local * tb], 0)]):
                    req.wait()
        else:
            pending, ops = [], []
            for r in range(1, self.world):  # one batched group of receives: all peers transfer concurrently
                if not counts[r]:
                    continue
                xy_t = torch.empty(counts[r] * 2, dtype=torch.int32, device=dev)
                buf = torch.empty(counts[r] * tb, dtype=torch.uint8, device=dev)
                pending.append((r, xy_t, buf))
                ops += [dist.P2POp(dist.irecv, xy_t, r), dist.P2POp(dist.irecv, buf, r)]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                if self.cuda:
                    torch.cuda.current_stream().synchronize()
            for r, xy_t, buf in pending:
                self.map.import_tiles(xy_t.cpu().numpy().reshape(-1, 2), buf.data_ptr(), self.cuda)
                received += counts[r]
        return received


def bench_main(args, rank, world, local_rank):
    """bench.py --gpus N (N>1), launched by torchrun: strong scaling of the BASELINE workload over N tile shards."""
    import json
    import pi_slam_fusion_b200.map2d as m2d
    import pi_slam_fusion_b200.synth as synth
    from bench import W, H, SEED, METRIC, UNIT, ClockSampler, workload_name

    # NCCL prints its version banner to stdout; the bench contract is ONE JSON line there -> park fd 1 on stderr
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode = args.mode
    typ = 3 if mode == "multiband" else 1
    n = args.frames
    seq = synth.Sequence(n, W, H, seed=SEED)
    dev = torch.device("cuda", local_rank)
    block = max(8, 128 // world)
    ids = ShardedMap2D.local_frame_ids(n, rank, world, block)
    frames = torch.from_numpy(np.stack([seq.frame(k) for k in ids])).to(dev)  # inputs resident in HBM, spread over the job's GPUs
    shard = default_shard(W, H, world, seq.poses, seq.camera)
    block = max(8, 128 // world)
    sm = ShardedMap2D(lambda t, **kw: m2d.Map2D.create(t, thread=False, **kw), typ, rank, world, device=local_rank, **shard)
    assert sm.prepare(seq.plane, seq.camera, seq.prepare_poses)

    def step():
        sm.map.reset()
        t0 = time.perf_counter()
        res = sm.feed_all_distributed(frames, seq.poses, W, H, block)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sm.gather_to_root()
        torch.cuda.synchronize()
        return res, t1 - t0, time.perf_counter() - t1

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sm.map.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_feed = t_gather = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, a, b = step()
        t_feed += a
        t_gather += b
    ev1.record()
    dist.barrier()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    ms = max(ev0.elapsed_time(ev1) / args.steps, 0.0)
    t = torch.tensor([max(ms, wall_ms), t_feed * 1e3 / args.steps, t_gather * 1e3 / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    launches = torch.tensor([float(sm.map.launch_count() - l0)], device=dev, dtype=torch.float64)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    clocks = sampler.result()
    fused = int((res == 0).sum())
    if rank == 0:
        ms_step = float(t[0])
        line = {"metric": METRIC, "value": fused * W * H / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "s16", "data": "synthetic",
                "config": {"workload": workload_name(mode, n), "mode": mode, "frames": n, "frames_fused": fused,
                           "parallelism": "%d tile shards (axis %d, span %d tiles); frames resident block-cyclically on all GPUs, all_gather per chunk (NCCL), final tile gather to rank 0"
                                          % (world, shard["shard_axis"], shard["shard_span"]),
                           "l2": "inputs %.2f GB per step > 126 MB L2" % (n * W * H * 3 / 1e9)},
                "clocks": clocks, "gpu_launches": int(launches.item() / args.steps),
                "breakdown_ms": {"allgather_plus_fuse": float(t[1]), "tile_gather": float(t[2])},
                "e2e": None, "roofline": None, "cpu_baseline": None}
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    dist.destroy_process_group()
