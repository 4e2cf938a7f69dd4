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


def strip_owner(a, span, world, origin=0):
    """Owner of absolute tile coordinate `a` along the shard axis (the rule of m2d_config.shard_* / m2d_set_shard)."""
    return int(np.floor((a - origin) / float(span))) % world


class DeliveryPlan:
    """Which frame pixels travel where when frames originate on the GPUs that captured them ("deliver
    frame + H + tile rect to owners ... P2P of just what maps into each owner's tiles").

    rects     [n,4] int: x0,y0,x1,y1 tile rect of every frame in ABSOLUTE tile coordinates (-1 = rejected), exactly as
              the sequential feeds will compute them: plan_rects() (m2d_plan_rects, a dry run that includes spreadMap);
    resident  per rank, the half-open range [lo,hi) of feed-order frame indices whose pixels start out on that rank.
    A rank needs the pixels of every frame whose rect (optionally grown by `margin` tiles) contains a tile it owns
    (a wrong plan is caught: m2d_feed_poses refuses a pose under which the shard owns a tile); it is handed the hull
    [a,b) of those frames (a serpentine survey cut into strips makes the needed frames contiguous; a hull that covers
    a few unneeded frames costs only their transfer — feeding a frame under which a shard owns nothing is a no-op).
    Frames outside the hull are fed as poses only (m2d_feed_poses), so every rank grows its grid like an unsharded run.
    """

    def __init__(self, rects, axis, span, world, resident, origin=0, margin=0):
        rects = np.asarray(rects, np.int64).reshape(-1, 4)
        self.n, self.world = len(rects), world
        self.resident = [tuple(r) for r in resident]
        assert len(self.resident) == world
        need = [[] for _ in range(world)]
        for k, (x0, y0, x1, y1) in enumerate(rects):
            if x1 <= x0 or y1 <= y0:      # rejected frame: nobody needs its pixels
                continue
            lo, hi = (x0, x1) if axis == 0 else (y0, y1)
            if hi - lo + 2 * margin >= span * world:
                owners = range(world)
            else:
                owners = {strip_owner(a, span, world, origin) for a in range(lo - margin, hi + margin)}
            for r in owners:
                need[r].append(k)
        self.needed = [np.array(v, np.int64) for v in need]
        self.hull = [(int(v[0]), int(v[-1]) + 1) if len(v) else (0, 0) for v in self.needed]
        # buffer of rank r = hull U resident range (its own frames live inside the buffer: no copy in the timed region)
        self.buffer = []
        for r in range(world):
            (a, b), (lo, hi) = self.hull[r], self.resident[r]
            if a == b:
                self.buffer.append((lo, hi))
            elif lo == hi:
                self.buffer.append((a, b))
            else:
                self.buffer.append((min(a, lo), max(b, hi)))
        # transfers (src, dst, lo, hi): the part of dst's hull that is resident on src
        self.transfers = []
        for d in range(world):
            a, b = self.hull[d]
            for s_ in range(world):
                if s_ == d:
                    continue
                lo, hi = max(a, self.resident[s_][0]), min(b, self.resident[s_][1])
                if lo < hi:
                    self.transfers.append((s_, d, lo, hi))
        covered = [sum(min(b, hi) - max(a, lo) for (lo, hi) in [self.resident[r]] + [(t[2], t[3]) for t in self.transfers if t[1] == r]
                       if min(b, hi) > max(a, lo)) for r, (a, b) in enumerate(self.hull)]
        for r, (a, b) in enumerate(self.hull):
            assert covered[r] == b - a, "resident ranges must partition the sequence"

    def frames_moved(self):
        return sum(hi - lo for (_, _, lo, hi) in self.transfers)


def even_split(n, world):
    """Contiguous, near-equal feed-order ranges: frames stay on the GPU whose flight lines produced them."""
    cuts = [(n * r) // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


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

    # ---- frames originate where they were captured; only halo frames move (weak scaling over survey area) ----
    def align_strips(self, poses, axis=None):
        """After prepare(): cut the tile range the survey will touch into `world` contiguous strips (one per rank)
        and install them with set_shard.  Every rank computes the same cut from the same poses.  Returns
        (rects, axis, span, origin) for DeliveryPlan."""
        rects = np.asarray(self.map.plan_rects(poses), np.int64).reshape(-1, 4)
        ok = rects[rects[:, 2] > rects[:, 0]]
        if len(ok) == 0:
            raise ValueError("no frame of the sequence is accepted")
        ext = [(int(ok[:, 0].min()), int(ok[:, 2].max())), (int(ok[:, 1].min()), int(ok[:, 3].max()))]
        if axis is None:
            # cut ACROSS the flight lines: along the axis on which consecutive frames move least (the cross-track axis), so
            # that a strip holds whole flight lines, its frames are contiguous in feed order and only the lines next to a
            # boundary are halo frames.  (Cutting along the track would make every rank need a piece of every line.)
            cx, cy = 0.5 * (ok[:, 0] + ok[:, 2]), 0.5 * (ok[:, 1] + ok[:, 3])
            if len(ok) > 1:
                axis = 0 if np.abs(np.diff(cx)).mean() <= np.abs(np.diff(cy)).mean() else 1
            else:
                axis = 0 if ext[0][1] - ext[0][0] >= ext[1][1] - ext[1][0] else 1
        lo, hi = ext[axis]
        span = max(1, -(-(hi - lo) // self.world))
        self.map.set_shard(self.rank, self.world, axis, span, lo)
        return rects, axis, span, lo

    def alloc_owned_buffer(self, plan, w, h):
        """uint8 [frames of plan.buffer[rank], h, w, 3] on this rank's device, and the view of it that holds the
        rank's resident frames (to be filled by the caller before the timed region)."""
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        b0, b1 = plan.buffer[self.rank]
        buf = torch.empty((max(b1 - b0, 1), h, w, 3), dtype=torch.uint8, device=dev)
        lo, hi = plan.resident[self.rank]
        return buf, buf[lo - b0:hi - b0]

    def feed_all_owned(self, plan, buf, poses, w, h):
        """One pass over the whole sequence.  Halo frames come by P2P from the ranks they are resident on (batched
        isend/irecv groups posted up front).  The rank feeds in global order -- poses only outside its hull, pixels
        inside, the whole hull in ONE m2d_feed_batch (large groups cull best) -- and the library's pixel-reading kernels
        wait on the device for the transfer (m2d_set_input_event), so it overlaps the bounds / weight / decide stages,
        which need poses only."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        assert n == plan.n
        b0, _ = plan.buffer[self.rank]

        def post(lower):
            ops = []
            for (src, dst, lo, hi) in plan.transfers:
                if (lo < plan.resident[dst][0]) != lower:
                    continue
                if src == self.rank:
                    ops.append(dist.P2POp(dist.isend, buf[lo - b0:hi - b0], dst))
                elif dst == self.rank:
                    ops.append(dist.P2POp(dist.irecv, buf[lo - b0:hi - b0], src))
            return dist.batch_isend_irecv(ops) if ops else []

        first, second = post(True), post(False)
        a, b = plan.hull[self.rank]
        res = np.zeros(n, np.int32)
        if a > 0:
            res[:a] = self.map.feed_poses(poses[:a])
        reqs = list(first) + list(second)
        for req in reqs:
            req.wait()               # NCCL: the CURRENT STREAM waits for the transfer, the host does not
        ev = None
        if reqs and self.cuda:
            if hasattr(self.map, "set_input_event"):
                # the library reads pixels on its own streams: its image kernels wait for this event, while bounds, weights
                # and winners (poses only) run at once -- the halo transfer hides behind them
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                self.map.set_input_event(ev.cuda_event)
            else:
                torch.cuda.current_stream().synchronize()
        if b > a:
            res[a:b] = self.map.feed_batch(buf[a - b0:].data_ptr(), b - a, w * h * 3, w, h, w * 3, poses[a:b], self.cuda)
        if b < n:
            res[b:] = self.map.feed_poses(poses[b:])
        self.map.sync()
        return res

    # ---- zero-copy frame delivery: frames stay where they were captured, neighbours sample them in place over NVLink ----
    def share_frames(self, plan, w, h):
        """CUDA only.  Put this rank's resident frames into a buffer the other processes can map (CUDA IPC), map the
        buffers of the ranks that hold frames of this rank's hull, and return (mine, ptrs): `mine` = uint8 tensor
        [resident frames, h, w, 3] to be filled by the caller, `ptrs` = one device address per frame of the sequence (0 where
        the rank needs no pixels).  The kernels of m2d_feed_batch_ptrs then read halo frames straight from the neighbour's HBM
        (P2P loads over NVLink/NVSwitch): no halo copy is made and only the px that are really sampled cross the link."""
        import pi_slam_fusion_b200.map2d as m2d
        fb = w * h * 3
        lo, hi = plan.resident[self.rank]
        self._frames = m2d.DeviceBuffer(self.device, max(hi - lo, 1) * fb)
        handles = [None] * self.world
        dist.all_gather_object(handles, self._frames.export())
        a, b = plan.hull[self.rank]
        ptrs = np.zeros(plan.n, np.uint64)
        self._mapped = []
        for r in range(self.world):
            rlo, rhi = plan.resident[r]
            k0, k1 = max(a, rlo), min(b, rhi)
            if r == self.rank:
                k0, k1 = rlo, rhi
                base = self._frames.ptr
            elif k1 > k0:
                base = m2d.ipc_open(self.device, handles[r])
                self._mapped.append(base)
            else:
                continue
            for k in range(k0, k1):
                ptrs[k] = base + (k - rlo) * fb
        mine = self._frames.tensor()[:(hi - lo) * fb].view(hi - lo, h, w, 3)
        return mine, ptrs

    def unshare_frames(self):
        import pi_slam_fusion_b200.map2d as m2d
        dist.barrier()   # nobody reads my frames any more
        for p in getattr(self, "_mapped", []):
            m2d.ipc_close(self.device, p)
        self._mapped = []
        if getattr(self, "_frames", None) is not None:
            self._frames.free()
            self._frames = None

    def feed_all_peer(self, plan, ptrs, poses, w, h, sync=True):
        """One pass over the whole sequence with share_frames() delivery: poses only outside the rank's hull, ONE
        m2d_feed_batch_ptrs inside it (own frames by local address, halo frames by their mapped peer address).  No
        collective, no copy: the only inter-GPU traffic are the kernels' own loads."""
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        n = len(poses)
        a, b = plan.hull[self.rank]
        res = np.zeros(n, np.int32)
        if a > 0:
            res[:a] = self.map.feed_poses(poses[:a])
        if b > a:
            res[a:b] = self.map.feed_batch_ptrs(ptrs[a:b], w, h, w * 3, poses[a:b], True)
        if b < n:
            res[b:] = self.map.feed_poses(poses[b:])
        if sync:   # (the frames are static inputs: back-to-back passes need no sync in between, the host may run ahead)
            self.map.sync()
        return res

    # ---- sharded save: every rank collapses its own strip (+ halo rows from its neighbours); nobody holds the whole map ----
    def save_sharded(self, axis, span, origin, levels=None, out=None, gather=False, chunk=None, sink=None):
        """The sharded half of Map2D::save (MultiBandMap2DCPU.cpp:779-847 / Map2DCPU.cpp:523-564) for the contiguous
        strips align_strips() installs (rank r owns absolute tile coordinates [origin + r*span, origin + (r+1)*span)
        along `axis`).  Collective.  Steps: all ranks agree on the global bbox of touched tiles (one tiny all_gather);
        neighbours swap the k tile rows next to their common boundary, raw (k = ceil((2^levels - 2) / 256): what a
        window edge that is not the mosaic's edge gets wrong reaches 2^levels - 2 level-0 px into the window); every
        rank collapses window = its strip + halo (m2d_get_image_rect) and keeps the crop = its strip.
        Returns (strip, rect_abs): strip = uint8 tensor [h, w, channels] on the rank's device (CUDA) or CPU (gloo), or None
        if the rank owns no row of the bbox, written into `out` (a flat uint8 tensor, grown as needed) if given;
        rect_abs = the strip's absolute tile rect.  gather=True additionally assembles the whole mosaic on rank 0
        (returned there instead of the strip; only sensible when it fits one device).  chunk=c collapses the strip c tile
        rows (columns) at a time and hands every piece to sink(tensor, rect_abs) -- a strip larger than the device's free
        memory (cfg4: 15 GB of BGRA per rank) is streamed to the host through one reusable buffer."""
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        levels = levels or getattr(self.map, "levels", 1)
        k = max(1, -(-((1 << levels) - 2) // 256)) if self.type == 3 else 0
        bb = self.map.tile_bbox()
        mine = torch.tensor(list(bb) + [1] if bb else [0, 0, 0, 0, 0], dtype=torch.int64, device=dev)
        allbb = [torch.zeros(5, dtype=torch.int64, device=dev) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(allbb, mine)
        else:
            allbb = [mine]
        allbb = np.array([t.cpu().numpy() for t in allbb])
        have = allbb[allbb[:, 4] == 1]
        if len(have) == 0:
            return None, None
        G = [int(have[:, 0].min()), int(have[:, 1].min()), int(have[:, 2].max()), int(have[:, 3].max())]
        g0, g1 = (G[0], G[2]) if axis == 0 else (G[1], G[3])
        if g0 < origin or g1 > origin + span * self.world or (self.world > 1 and span < k):
            raise RuntimeError("save_sharded needs the contiguous strips of align_strips(); use gather_to_root() for other layouts")

        def interval(r):
            return max(origin + r * span, g0), min(origin + (r + 1) * span, g1)

        def rect(a, b):   # tile rect of rows/columns [a, b) along the shard axis, full extent across
            return (a, G[1], b, G[3]) if axis == 0 else (G[0], a, G[2], b)

        lo, hi = interval(self.rank)
        tb = self.map.tile_bytes()
        halos = []
        if k and self.world > 1:
            # what I send: my first k rows to the rank below, my last k rows to the rank above (if they own anything)
            sends, recvs = [], []
            for nb, (a, b) in ((self.rank - 1, (lo, min(lo + k, hi))), (self.rank + 1, (max(hi - k, lo), hi))):
                if nb < 0 or nb >= self.world or hi <= lo:
                    continue
                nlo, nhi = interval(nb)
                if nhi <= nlo:
                    continue   # NOTE: an empty neighbour strip in the middle of the bbox is not handled (align_strips never makes one)
                n = self.map.export_tiles_rect(rect(a, b), 0, 0, self.cuda) if b > a else 0
                buf = torch.empty(max(n, 1) * tb, dtype=torch.uint8, device=dev)
                xy = self.map.export_tiles_rect(rect(a, b), buf.data_ptr(), n, self.cuda) if n else np.zeros((0, 2), np.int32)
                sends.append((nb, n, torch.from_numpy(np.ascontiguousarray(xy.reshape(-1))).to(dev), buf))
            # counts first (neighbours only), then the tiles
            cnt_out = {nb: torch.tensor([n], dtype=torch.int64, device=dev) for nb, n, _, _ in sends}
            cnt_in = {nb: torch.zeros(1, dtype=torch.int64, device=dev) for nb, _, _, _ in sends}
            ops = [dist.P2POp(dist.isend, cnt_out[nb], nb) for nb in cnt_out] + [dist.P2POp(dist.irecv, cnt_in[nb], nb) for nb in cnt_in]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            ops = []
            for nb, n, xy_t, buf in sends:
                if n:
                    ops += [dist.P2POp(dist.isend, xy_t, nb), dist.P2POp(dist.isend, buf[:n * tb], nb)]
                m = int(cnt_in[nb].item())
                if m:
                    xy_r = torch.empty(m * 2, dtype=torch.int32, device=dev)
                    buf_r = torch.empty(m * tb, dtype=torch.uint8, device=dev)
                    recvs.append((xy_r, buf_r))
                    ops += [dist.P2POp(dist.irecv, xy_r, nb), dist.P2POp(dist.irecv, buf_r, nb)]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                if self.cuda:
                    torch.cuda.current_stream().synchronize()
            for xy_r, buf_r in recvs:
                self.map.import_tiles(xy_r.cpu().numpy().reshape(-1, 2), buf_r.data_ptr(), self.cuda)
            halos = [rect(max(lo - k, g0), lo), rect(hi, min(hi + k, g1))]
        strip, srect = None, None
        if hi > lo:
            srect = rect(lo, hi)
            cn = 3 if self.type == 3 else 4
            step = chunk if (chunk and sink) else (hi - lo)
            for c0 in range(lo, hi, step):
                c1 = min(c0 + step, hi)
                win, crop = rect(max(c0 - k, g0), min(c1 + k, g1)), rect(c0, c1)
                hpx, wpx = (crop[3] - crop[1]) * 256, (crop[2] - crop[0]) * 256
                need = hpx * wpx * cn
                if out is None:
                    out = getattr(self, "_save_buf", None)
                if out is None or out.numel() < need:
                    out = self._save_buf = torch.empty(need, dtype=torch.uint8, device=dev)   # cached: cudaMalloc of GBs is slow
                self.map.get_image_rect(win, crop, out.data_ptr(), self.cuda)
                strip = out[:need].view(hpx, wpx, cn)
                if sink:
                    sink(strip, crop)
            if chunk and sink:
                strip = None
        for r_ in halos:
            if r_[2] > r_[0] and r_[3] > r_[1]:
                self.map.drop_tiles_rect(r_)
        if not gather:
            return strip, srect
        # optional: the whole mosaic on rank 0 (192 KiB per tile instead of 874 KB of raw state, already collapsed)
        shape = torch.tensor(list(strip.shape) if strip is not None else [0, 0, 0], dtype=torch.int64, device=dev)
        shapes = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(shapes, shape)
        else:
            shapes = [shape]
        shapes = [tuple(int(v) for v in t.tolist()) for t in shapes]
        if self.rank != 0:
            if strip is not None:
                dist.send(strip.contiguous().view(-1), 0)
            return None, G
        parts = []
        for r in range(self.world):
            if shapes[r][0] == 0:
                continue
            if r == 0:
                parts.append(strip)
            else:
                t = torch.empty(shapes[r], dtype=torch.uint8, device=dev)
                dist.recv(t.view(-1), r)
                parts.append(t)
        return torch.cat(parts, dim=1 if axis == 0 else 0), G

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
    """bench.py --gpus N (N>1), launched by torchrun.  Default: weak scaling over survey area (bench_weak);
    --scaling strong keeps the BASELINE workload fixed and cuts it into N tile shards (bench_strong)."""
    if getattr(args, "scaling", "weak") == "strong":
        return bench_strong(args, rank, world, local_rank)
    return bench_weak(args, rank, world, local_rank)


def _sha(t):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy() if hasattr(t, "cpu") else t).tobytes()).hexdigest()[:16]


def _strip_run(args, rank, world, local_rank, mode, seq, label, steps, warm, check_unsharded, want_e2e, B, m2d, want_sha=False):
    """One tile-sharded job on `world` GPUs: contiguous strips of tiles (align_strips), frames resident in the HBM of the GPU
    whose flight lines produced them, halo frames by NCCL P2P inside the timed region, poses to all ranks.  Timed: a step =
    reset + exchange + fuse (max over ranks, device events and wall clock); then the sharded save on its own (halo tile
    rows between neighbours, per-rank collapse of its own strip, D2H of the strip into pinned host memory).
    Returns the result dict on rank 0 (None elsewhere)."""
    BB = B
    W, H, n = seq.w, seq.h, seq.n
    typ = 3 if mode == "multiband" else 1
    dev = torch.device("cuda", local_rank)
    sm = ShardedMap2D(lambda t, **kw: m2d.Map2D.create(t, thread=False, **kw), typ, rank, world, device=local_rank,
                      shard_axis=0, shard_span=4, batch_frames=args.batch)
    assert sm.prepare(seq.plane, seq.camera, seq.prepare_poses)
    rects, axis, span, origin = sm.align_strips(seq.poses)
    plan = DeliveryPlan(rects, axis, span, world, even_split(n, world), origin)
    peer = getattr(args, "halo", "peer") == "peer"
    if peer:   # frames stay on the GPU that "captured" them; neighbours sample halo frames in place over NVLink (CUDA IPC)
        buf = None
        mine, ptrs = sm.share_frames(plan, W, H)
    else:      # halo frames are copied by NCCL P2P inside every step
        buf, mine = sm.alloc_owned_buffer(plan, W, H)
    lo, hi = plan.resident[rank]
    for c0 in range(lo, hi, 64):   # inputs resident in HBM before the timed region
        c1 = min(c0 + 64, hi)
        mine[c0 - lo:c1 - lo].copy_(BB.device_frames(torch, seq, c0, c1, dev))
    torch.cuda.synchronize()
    dist.barrier()

    def step(sync=True):
        sm.map.reset()
        if peer:
            return sm.feed_all_peer(plan, ptrs, seq.poses, W, H, sync=sync)
        return sm.feed_all_owned(plan, buf, seq.poses, W, H)

    for _ in range(warm):
        res = step()
    sm.map.sync()
    sampler = B.ClockSampler(local_rank)
    sampler.start()
    l0 = sm.map.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = step(sync=False)   # like the 1-GPU bench: passes back to back, the host prepares pass i+1 while pass i is being fused
    sm.map.sync()                # the library's streams have drained
    ev1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    dist.barrier()
    torch.cuda.synchronize()
    ms = max(ev0.elapsed_time(ev1) / steps, wall_ms)
    launches = float(sm.map.launch_count() - l0) / steps
    clocks = sampler.result()

    # ---- sharded save: warm once (NCCL connections, collapse buffers), then time.  The strip is collapsed 16 tile rows
    # at a time and streamed to the host through one reusable pinned buffer (a rank's strip may exceed its free HBM: cfg4)
    sink_state = {"host": None, "bytes": 0}

    def sink(t, rect_abs):
        nb = int(t.numel())
        if sink_state["host"] is None or sink_state["host"].numel() < nb:
            sink_state["host"] = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        sink_state["host"][:nb].copy_(t.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the device buffer is reused by the next piece
        sink_state["bytes"] += nb

    def save():
        sink_state["bytes"] = 0
        sm.save_sharded(axis, span, origin, chunk=16, sink=sink)

    save()
    nbytes = sink_state["bytes"]

    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    save()
    torch.cuda.synchronize()
    save_ms = (time.perf_counter() - t0) * 1e3

    # ---- e2e: host buffers -> H2D of the rank's own frames, halo exchange, fuse, sharded save, strip D2H on every rank
    e2e_ms = None
    if want_e2e:
        host_in = torch.empty((hi - lo, H, W, 3), dtype=torch.uint8, pin_memory=True)
        host_in.copy_(mine)

        def step_e2e():
            sm.map.reset()
            mine.copy_(host_in, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if peer:
                dist.barrier()        # the neighbours' frames have landed too before anybody samples them ...
                sm.feed_all_peer(plan, ptrs, seq.poses, W, H)
                dist.barrier()        # ... and nobody re-uploads while a neighbour still reads
            else:
                sm.feed_all_owned(plan, buf, seq.poses, W, H)
            save()

        step_e2e()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            step_e2e()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
        del host_in

    # ---- correctness inside the run: the strips, assembled on rank 0, against an UNSHARDED run of the same survey on rank 0
    parity = None
    if check_unsharded:
        step()
        mosaic, gbox = sm.save_sharded(axis, span, origin, gather=True)
        if rank == 0:
            ref = m2d.Map2D.create(typ, thread=False, device=local_rank, batch_frames=args.batch)
            assert ref.prepare(seq.plane, seq.camera, seq.prepare_poses)
            chunk = max(1, min(n, int(6e9 // (W * H * 3))))
            for c0 in range(0, n, chunk):
                fr = BB.device_frames(torch, seq, c0, min(c0 + chunk, n), dev)
                ref.feed_batch(fr.data_ptr(), fr.shape[0], W * H * 3, W, H, W * 3, seq.poses[c0:c0 + fr.shape[0]], True)
                ref.sync()
                del fr
            img, org = ref.get_image()
            same = bool(tuple(img.shape) == tuple(mosaic.shape) and np.array_equal(img, mosaic.cpu().numpy()))
            parity = {"checked": "mosaic of the sharded save (strips of all ranks) vs an unsharded run of the same %d frames on rank 0" % n,
                      "identical": same, "sha256_sharded": _sha(mosaic), "sha256_unsharded": _sha(img), "mosaic_px": int(img.shape[0] * img.shape[1])}
            ref.close()
            del img
        del mosaic
    mosaic_sha = None
    if want_sha and not check_unsharded:   # the assembled mosaic's digest, comparable across runs with different N
        mosaic, gbox = sm.save_sharded(axis, span, origin, gather=True)
        if rank == 0:
            mosaic_sha = _sha(mosaic)
        del mosaic
    t = torch.tensor([ms, save_ms, e2e_ms or 0.0, launches], device=dev, dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    hull = torch.tensor([plan.hull[rank][1] - plan.hull[rank][0], float(nbytes)], device=dev, dtype=torch.float64)
    hmax = hull.clone()
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hull, op=dist.ReduceOp.SUM)
    fused = int((res == 0).sum())
    tiles = torch.tensor([float(sm.map.tile_count())], device=dev, dtype=torch.float64)
    dist.all_reduce(tiles, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        ms_step, sv = float(tmax[0]), float(tmax[1])
        px = fused * W * H
        out = {"label": label, "mode": mode, "value": px / (ms_step * 1e-3) / 1e6, "unit": B.UNIT, "ms_per_step": ms_step, "frames": n, "frames_fused": fused,
               "gpu_launches": int(float(t[3])), "clocks": clocks, "tiles": int(tiles.item()),
               "parallelism": "%d contiguous strips of %d tiles along axis %d (tile ownership); frames resident on the GPU of their own flight lines, "
                              "%d halo frames %s, poses to all ranks; largest rank feeds %d of %d frames"
                              % (world, span, axis, plan.frames_moved(),
                                 "sampled IN PLACE from the neighbour's HBM over NVLink (CUDA IPC peer loads inside the fusion kernels, no copies)" if peer
                                 else "exchanged by NCCL P2P per step (inside the timed region)", int(hmax[0].item()), n),
               "save": {"what": "sharded save: neighbours swap one raw tile row per boundary, every rank collapses ITS strip and copies it to pinned host memory; no rank holds the whole map",
                        "ms": sv, "mosaic_bytes": int(hull[1].item()), "value_incl_save": px / ((ms_step + sv) * 1e-3) / 1e6},
               "parity": parity}
        if mosaic_sha:
            out["mosaic_sha256"] = mosaic_sha
        if e2e_ms is not None:
            out["e2e"] = {"value": px / (float(tmax[2]) * 1e-3) / 1e6, "unit": B.UNIT, "h2d_bytes_per_step": n * W * H * 3,
                          "d2h_bytes_per_step": int(hull[1].item()), "ms_per_step": float(tmax[2]),
                          "what": "per rank: H2D of its own pinned host frames, halo exchange, m2d_feed_batch, sharded save (m2d_get_image_rect of its strip) + D2H of the strip"}
    sm.map.close()
    del buf, mine, sink_state
    if peer:
        sm.unshare_frames()
    sm._save_buf = None
    torch.cuda.empty_cache()
    return out


def bench_weak(args, rank, world, local_rank):
    """Headline: N x the BASELINE cfg2 survey (same flight-line length, N x the lines), cut into N strips of tiles -- per-GPU
    work is fixed (weak scaling).  The same line carries the weighted mode on the same survey, and cfg3 (BASELINE configs[2]:
    the 1000-frame 4000x3000 multi-band survey) as ONE fixed job cut into N strips (strong scaling)."""
    import json
    import pi_slam_fusion_b200.map2d as m2d
    import pi_slam_fusion_b200.synth as synth
    import bench as B

    saved_stdout = os.dup(1)   # NCCL prints its banner to stdout; the contract is ONE JSON line there
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    head_cfg = B.CFG2 if args.mode == "multiband" else B.CFG1
    other_mode = "weighted" if args.mode == "multiband" else "multiband"
    mode, per_gpu, W, H, seed, _ = head_cfg
    per_gpu = args.frames or B.CFG2[1]          # the weak-scaled survey is cfg2-shaped in both modes (500 frames per GPU)
    W, H, seed = B.CFG2[2], B.CFG2[3], B.CFG2[4]
    if args.size:
        W, H = (int(v) for v in args.size.lower().split("x"))
    n = per_gpu * world
    fpl = synth.frames_per_line(per_gpu, W, H)
    seq = synth.Sequence(n, W, H, seed=seed, fpl=fpl)
    warm = max(args.warmup, 3)
    head = _strip_run(args, rank, world, local_rank, mode, seq, "headline", args.steps, warm, True, not args.no_e2e, B, m2d)
    other = cfg3 = cfg4 = None
    if not args.only:
        other = _strip_run(args, rank, world, local_rank, other_mode, seq, other_mode, max(3, args.steps // 2), 3, True, not args.no_e2e, B, m2d)
        if args.cfg3_frames > 0:
            _, n3, w3, h3, s3, _ = B.CFG3
            n3 = args.cfg3_frames
            seq3 = synth.Sequence(n3, w3, h3, seed=s3)
            try:
                cfg3 = _strip_run(args, rank, world, local_rank, "multiband", seq3, "cfg3", max(2, min(args.steps, 5)), 2, False, False, B, m2d, want_sha=True)
            except Exception as e:   # (an error that hits every rank alike, e.g. out of memory: the rest of the line is still worth printing)
                cfg3 = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
        n4 = args.cfg4_frames if args.cfg4_frames >= 0 else (B.CFG4[1] if world >= 8 else 0)
        if n4 > 0:
            # BASELINE configs[3]: weighted fusion of a 5000-frame 4000x3000 survey with 30 % / 30 % overlap: ~470k tiles =
            # ~120 GB of BGRA map state, more than one GPU's tile budget; sharded, saved strip by strip, never gathered
            _, _, w4, h4, s4, _ = B.CFG4
            seq4 = synth.Sequence(n4, w4, h4, seed=s4, along=0.7, cross=0.7)
            try:
                cfg4 = _strip_run(args, rank, world, local_rank, "weighted", seq4, "cfg4", 3, 2, False, False, B, m2d)
            except Exception as e:
                cfg4 = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
    if rank == 0:
        cfg = B.config_weak(mode, per_gpu, W, H, seed, world)
        line = {"metric": B.METRIC, "value": head["value"], "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s16" if mode == "multiband" else "u8", "data": "synthetic", "config": cfg,
                "run": {"frames_fused": head["frames_fused"], "parallelism": head["parallelism"], "tiles": head["tiles"]},
                "clocks": head["clocks"], "gpu_launches": head["gpu_launches"],
                "breakdown_ms": {"exchange_plus_fuse": head["ms_per_step"], "sharded_save": head["save"]["ms"],
                                 "value_incl_save": head["save"]["value_incl_save"]},
                "save": head["save"], "e2e": head.get("e2e"), "parity": head["parity"], "roofline": None, "cpu_baseline": None}
        if other:
            line[other_mode] = other
        if cfg3 and "error" in cfg3:
            line["cfg3"] = cfg3
        elif cfg3:
            cfg3["workload"] = B.workload_name("multiband", seq3.n, seq3.w, seq3.h, B.CFG3[4])
            cfg3["scaling"] = "strong"
            cfg3["n_gpus"] = world
            line["cfg3"] = cfg3
        if cfg4 and "error" in cfg4:
            line["cfg4"] = cfg4
        elif cfg4:
            cfg4["workload"] = "cfg4: Map2DCPU weighted fusion, %d synthetic %dx%d frames, 30 %%/30 %% overlap (seed %d): map state %.1f GB across %d GPUs" % (
                seq4.n, seq4.w, seq4.h, B.CFG4[4], cfg4["tiles"] * 262144 / 1e9, world)
            cfg4["scaling"] = "strong"
            cfg4["n_gpus"] = world
            line["cfg4"] = cfg4
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    dist.destroy_process_group()


def bench_strong(args, rank, world, local_rank):
    """Strong scaling of the BASELINE workload over N tile shards (frames block-cyclic over the GPUs, all_gather)."""
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
