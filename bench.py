#!/usr/bin/env python
"""bench.py — fused input Mpix/s of Map2D::feed() (BASELINE.json metric: "weighted & multi-band") on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode multiband|weighted]

ONE JSON line.  Headline (`value`, `e2e`, `roofline`, `cpu_baseline`) = BASELINE.json configs[1] (cfg2: MultiBandMap2DCPU,
5 bands, 500 nadir 1280x720 frames, seed 2); a "step" = reset the map, feed the whole survey.  The same line carries
  weighted        configs[0] (cfg1: Map2DCPU weighted fusion, 100 x 1280x720, seed 1): value, e2e, roofline, cpu_baseline
  stream_latency  configs[4] (cfg5: 2000 synchronous feed() calls, 1920x1080, pose jitter, H2D included): p50 / p99 ms
  cfg3            configs[2] (MultiBandMap2DCPU on 4000x3000 frames, 1000-frame survey): value on this many GPUs (the SAME
                  job at every N: strong scaling, tile-sharded for N > 1)
  cfg4            configs[3] (weighted fusion of a 5000-frame 4000x3000 survey, ~120 GB of map state): at 8 GPUs only
  parity          the mosaic of the fed prefix, bit for bit against the CPU oracle (sha256 of both)
Per measured mode:
  value     whole-job Mpix/s with the frames already resident in HBM (one m2d_feed_batch on device pointers per step)
  e2e       the same through the public host-buffer API: frames in pinned host memory, H2D copies inside the timed
            region, plus the collapsed mosaic (Map2D::save in memory) read back to the host every step
  roofline  dominant kernel: stage-algorithmic bytes per launch / CUDA-event launch duration (measured live via
            m2d_profile) against MEASURED_PEAKS.json; "path" = SURVEY §8(d) whole-frame algorithmic bytes / step time
  cpu_baseline  the CPU oracle (a port of the reference recipe), 1 thread, on a bounded sample of the same frames
`--impl reference` times the oracle with all host threads instead (the reference cannot be compiled here),
on the same config; under torchrun only rank 0 works.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pi_slam_fusion_b200.synth as synth  # noqa: E402

METRIC = "fused input Mpix/s"
UNIT = "Mpix/s"
# BASELINE.json configs: (mode, frames, W, H, seed, jitter)
CFG1 = ("weighted", 100, 1280, 720, 1, False)
CFG2 = ("multiband", 500, 1280, 720, 2, False)
CFG3 = ("multiband", 1000, 4000, 3000, 3, False)
CFG4 = ("weighted", 5000, 4000, 3000, 4, False)
CFG5 = (None, 2000, 1920, 1080, 5, True)


def workload_name(mode, n, w, h, seed):
    if mode == "multiband":
        tag = "cfg2" if (n, w, h, seed) == CFG2[1:5] else ("cfg3" if (w, h) == (4000, 3000) else "cfg2-shaped")
        return "%s: MultiBandMap2DCPU 5-band Laplacian blend, %d synthetic %dx%d BGR nadir frames (serpentine, seed %d)" % (tag, n, w, h, seed)
    tag = "cfg1" if (n, w, h, seed) == CFG1[1:5] else "cfg1-shaped"
    return "%s: Map2DCPU weighted fusion, %d synthetic %dx%d BGR nadir frames (serpentine, seed %d)" % (tag, n, w, h, seed)


def config_of(mode, n, w, h, seed, n_gpus=1):
    """The `config` object of the line: names the workload only (identical for the product arm and the reference arm)."""
    return {"workload": workload_name(mode, n, w, h, seed), "mode": mode, "frames": n, "frame": [w, h],
            "bands": 5 if mode == "multiband" else 0, "n_gpus": n_gpus,
            "l2": "inputs %.2f GB per step > 126 MB L2 (no flush needed)" % (n * w * h * 3 / 1e9)}


def config_weak(mode, per_gpu, w, h, seed, world):
    """`config` of the N > 1 line (both arms): the weak-scaled survey = N x the cfg2-shaped survey, same flight-line length."""
    n = per_gpu * world
    fpl = synth.frames_per_line(per_gpu, w, h)
    cfg = config_of(mode, n, w, h, seed, world)
    cfg["frames_per_gpu"] = per_gpu
    cfg["workload"] += " x %d GPUs: %d frames, %d flight lines of %d" % (world, n, -(-n // fpl), fpl)
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def algorithmic_bytes(mode, stats, levels):
    """compulsory HBM traffic of the sequential per-frame semantics (no intermediates)."""
    b = 3 * stats["input_px"]
    if mode == "multiband":
        for l in range(levels):
            nonfresh = stats["region_px"][l] - stats["fresh_px"][l]
            b += 4 * nonfresh + 10 * (stats["fresh_px"][l] + stats["win_px"][l])
    else:
        b += 4 * stats["footprint_px"] + 4 * stats["win_px"][0] + 4 * stats["fresh_px"][0]
    return b


def upper_bound_bytes(mode, stats, levels):
    """SURVEY.md §8(d) upper-bound variant: 11 B per input px (weighted); 3*W*H + (4/3)*14*D per frame (multi-band)."""
    if mode == "multiband":
        return 3 * stats["input_px"] + (4.0 / 3.0) * 14 * stats["region_px"][0]
    return 11 * stats["input_px"]


def stage_bytes(mode, stats, need, levels):
    """Per kernel class, per STEP: bytes the stage must move given ITS inputs/outputs (scratch pyramid counted at the
    reference's element sizes: 6 B int16x3 + 4 B f32 per px).  Divided by the class's launch count -> per launch.
    `need`: need_px / needw_px counted by the library during the profiled pass (px of the frames' Gaussian / weight level l
    that were actually computed)."""
    out = {}
    if mode == "weighted":
        out["weighted_fuse"] = algorithmic_bytes(mode, stats, levels)
        return out
    D = [stats["region_px"][l] for l in range(levels)]
    out["mb_warp"] = 3 * stats["input_px"] + 10 * D[0]
    out["mb_pyrdown"] = sum(10 * D[l] + 10 * D[l + 1] for l in range(levels - 2))
    out["mb_pyrtail"] = sum(10 * D[l] + 10 * D[l + 1] for l in range(max(levels - 2, 0), levels - 1))
    sel = dec = lap = 0.0
    N, NW = need["need_px"], need["needw_px"]
    for l in range(levels):
        nonfresh = stats["region_px"][l] - stats["fresh_px"][l]
        written = stats["fresh_px"][l] + stats["win_px"][l]
        sel += 4 * D[l] + 4 * nonfresh + written * (6 + 10 + (6 / 4 if l + 1 < levels else 0))
        lap += written * (6 + 6 + (6 / 4 if l + 1 < levels else 0))  # winner's G_l, its share of G_{l+1}, the Laplacian out
    out["mb_select"] = sel
    # weights-first pipeline (default): weights and image only in the cells that matter (needw_px / need_px, counted
    # exactly by the library); f32 weights at 4 B, image stages at the reference's 6 B int16x3 per px, source at 3 B
    out["mbw_warp"] = 4 * NW[0]
    out["mbw_pyramid"] = sum(4 * NW[l] + 4 * NW[l + 1] for l in range(levels - 1))
    out["mbs_warp"] = (3 + 6) * N[0]
    out["mbs_pyramid"] = sum(6 * N[l] + 6 * N[l + 1] for l in range(levels - 1))
    out["mbs_lap"] = lap
    return out


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture of
    this same workload (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    return t.get(kernel, {}).get("dram_bytes_per_launch")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle, all host threads
# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the CPU oracle (port of the reference recipe; the reference itself needs OpenCV 2.4 C++ / Qt / GL and
    cannot be compiled here), all host threads, on the product arm's config.  A step feeds a bounded sample of the
    workload (the whole 500-frame survey at N = 1: ~6 s with 16 threads)."""
    if rank != 0:
        return
    from oracle import oracle as O
    mode, n, w, h, seed, _ = (CFG2 if args.mode == "multiband" else CFG1)
    if world > 1:                                  # N > 1: the weak-scaled survey of the product arm (cfg2-shaped in both modes)
        n, w, h, seed = CFG2[1:5]
    n = args.frames or n
    if args.size:
        w, h = (int(v) for v in args.size.lower().split("x"))
    typ = 3 if mode == "multiband" else 1
    threads = os.cpu_count() or 1
    O.set_threads(threads)
    n_job = n * world if world > 1 else n
    fpl = synth.frames_per_line(n, w, h) if world > 1 else None
    seq = synth.Sequence(n_job, w, h, seed=seed, fpl=fpl)
    sample = min(n_job, args.ref_frames if args.ref_frames > 0 else max(4, int(500 * 1280 * 720 / (w * h))))
    frames = [seq.frame(k) for k in range(sample)]

    def step():
        o = O.OracleMap2D.create(typ)
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(sample):
            o.feed(frames[k], seq.poses[k])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample * w * h / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16" if mode == "multiband" else "u8", "data": "synthetic",
            "config": config_weak(mode, n, w, h, seed, world) if world > 1 else config_of(mode, n_job, w, h, seed, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "first %d frames of the %d-frame workload per step (oracle/map2d_oracle.cpp, OpenMP, %d threads)" % (sample, n_job, threads)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# one mode on one GPU: value, e2e, roofline, cpu_baseline, parity
# ---------------------------------------------------------------------------------------------------------------
def measure_mode(args, m2d, torch, mode, n, w, h, seed, local_rank, stream, want_clocks=True):
    typ = 3 if mode == "multiband" else 1
    seq = synth.Sequence(n, w, h, seed=seed)
    frame_bytes = w * h * 3
    host, host_ptr = m2d.pinned_empty((n, h, w, 3))
    for k in range(n):
        host[k] = seq.frame(k)
    dev = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    dev.copy_(torch.from_numpy(host))
    torch.cuda.synchronize()

    m = m2d.Map2D.create(typ, thread=False, device=local_rank, batch_frames=args.batch)
    m.set_stream(stream.cuda_stream)
    assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)

    def step_device():
        m.reset()
        return m.feed_batch(dev.data_ptr(), n, frame_bytes, w, h, w * 3, seq.poses, True)

    # ---- exact algorithmic bytes from one counted pass (separate handle, outside every timed region)
    mc = m2d.Map2D.create(typ, thread=False, device=local_rank, collect_stats=1, batch_frames=args.batch)
    assert mc.prepare(seq.plane, seq.camera, seq.prepare_poses)
    mc.feed_batch(dev.data_ptr(), n, frame_bytes, w, h, w * 3, seq.poses, True)
    mc.sync()
    stats = mc.stats()
    levels = mc.levels
    mc.close()
    fused = stats["frames_fused"]

    # ---- value: frames resident in HBM
    for _ in range(args.warmup):
        step_device()
    m.sync()
    sampler = ClockSampler(local_rank)
    if want_clocks:
        sampler.start()
    l0 = m.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    m.sync()
    torch.cuda.synchronize()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.result() if want_clocks else None
    launches = (m.launch_count() - l0) // args.steps
    value = fused * w * h / (ms_step * 1e-3) / 1e6

    # ---- live per-kernel timing (CUDA events around every launch on the launching stream)
    m.reset()
    before = m.stats()
    m.profile(True)
    m.feed_batch(dev.data_ptr(), n, frame_bytes, w, h, w * 3, seq.poses, True)
    kt = m.kernel_times()
    m.profile(False)
    after = m.stats()
    need = {k: [a - b for a, b in zip(after[k], before[k])] for k in ("need_px", "needw_px")}
    peak, peak_src = peaks()
    sb = stage_bytes(mode, stats, need, levels)
    total_kernel_ms = sum(v[0] for v in kt.values())
    dom = max(kt.items(), key=lambda kv: kv[1][0])[0]
    per_kernel = {}
    groups = max([c for k, (_, c) in kt.items() if k in ("mbs_decide", "mb_select", "weighted_fuse")] or [1])  # launch groups per step
    for k, (ms, cnt) in kt.items():
        per_launch_us = ms / cnt * 1e3
        bpl = sb[k] / cnt if k in sb else None  # the profiled pass is exactly one step
        ach = bpl / (ms / cnt * 1e-3) / 1e9 if bpl else None
        per_kernel[k] = {"launches": cnt, "avg_us": round(per_launch_us, 3), "share": round(ms / total_kernel_ms, 4),
                         "bytes_per_launch": bpl, "achieved_gbs": ach, "frames_per_launch": round(fused / groups, 2)}
    path_bytes = algorithmic_bytes(mode, stats, levels)
    path_ach = path_bytes / (ms_step * 1e-3) / 1e9
    ub_bytes = upper_bound_bytes(mode, stats, levels)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": (per_kernel[dom]["achieved_gbs"] or 0) / peak, "traffic": ncu_traffic(dom), "peak_source": peak_src,
                "kernels": per_kernel,
                "need_px_fraction": {"image": [round(need["need_px"][l] / max(stats["region_px"][l], 1), 4) for l in range(levels)],
                                     "weight": [round(need["needw_px"][l] / max(stats["region_px"][l], 1), 4) for l in range(levels)]} if mode == "multiband" else None,
                "path": {"bytes_per_step": path_bytes, "bytes_per_input_px": path_bytes / (fused * w * h),
                         "achieved": path_ach, "frac": path_ach / peak,
                         "upper_bound_bytes_per_input_px": ub_bytes / (fused * w * h), "frac_upper_bound": ub_bytes / (ms_step * 1e-3) / 1e9 / peak}}

    # ---- e2e: host buffers through the public API, H2D inside, mosaic read back
    e2e = None
    if not args.no_e2e:
        me = m2d.Map2D.create(typ, thread=False, device=local_rank, batch_frames=args.batch)
        me.set_stream(stream.cuda_stream)
        assert me.prepare(seq.plane, seq.camera, seq.prepare_poses)
        me.feed_batch(host_ptr, n, frame_bytes, w, h, w * 3, seq.poses, False)
        img, _ = me.get_image()
        out_bytes = img.nbytes
        del img
        out_pinned, out_ptr = m2d.pinned_empty((out_bytes,))

        zc = bool(getattr(args, "zero_copy", False))   # experiment: the kernels sample the pinned host frames in place over PCIe

        def step_e2e():
            me.reset()
            me.feed_batch(host_ptr, n, frame_bytes, w, h, w * 3, seq.poses, zc)
            return me.get_image(out=out_pinned)

        def step_e2e_split():   # the same, with a sync between the two halves: where the time goes (not the reported number)
            me.reset()
            t0 = time.perf_counter()
            me.feed_batch(host_ptr, n, frame_bytes, w, h, w * 3, seq.poses, zc)
            me.sync()
            t1 = time.perf_counter()
            me.get_image(out=out_pinned)
            return (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3

        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        reps = max(2, min(args.steps, 3))
        for _ in range(reps):
            step_e2e()
        e1.record(stream)
        me.sync()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps
        ms_e2e = max(e0.elapsed_time(e1) / reps, wall * 1e3)
        feed_ms, save_ms = step_e2e_split()
        e2e_sha = sha(out_pinned[:out_bytes])
        e2e = {"mosaic_sha256": e2e_sha, "frames_passed_as_device_pointers": zc,
               "host_frames": "pinned; not staged whole: the library pulls the needed 256-byte chunks (multi-band) / samples in place (weighted); M2D_ZEROCOPY=%s" % os.environ.get("M2D_ZEROCOPY", "1"),"value": fused * w * h / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": n * frame_bytes,
               "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_e2e,
               "breakdown_ms": {"feed_batch_from_host": feed_ms, "collapse_and_d2h": save_ms,
                                "h2d_gbs": n * frame_bytes / (feed_ms * 1e-3) / 1e9},
               "what": "m2d_feed_batch(host pinned frames) + m2d_get_image (collapse + D2H of the mosaic)"}
        me.close()
        m2d.free_pinned(out_ptr)

    # ---- CPU baseline (the oracle, one thread, bounded sample) and parity of the same prefix
    cpu = parity = None
    if not args.no_cpu:
        from oracle import oracle as O
        O.set_threads(1)
        o = O.OracleMap2D.create(typ)
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        ns = min(args.cpu_frames, n)
        t0 = time.perf_counter()
        for k in range(ns):
            o.feed(host[k], seq.poses[k])
        dt = time.perf_counter() - t0
        cpu = {"value": ns * w * h / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "first %d frames of the workload, oracle/map2d_oracle.cpp single thread (%.1f s)" % (ns, dt),
               "host_cores": os.cpu_count()}
        # parity: the same prefix through ONE m2d_feed_batch (default grouping, culling on), mosaic against the oracle's
        mp = m2d.Map2D.create(typ, thread=False, device=local_rank, batch_frames=args.batch)
        assert mp.prepare(seq.plane, seq.camera, seq.prepare_poses)
        mp.feed_batch(dev.data_ptr(), ns, frame_bytes, w, h, w * 3, seq.poses[:ns], True)
        gi, go = mp.get_image()
        oi, oo = o.get_image()
        same = bool(go == oo and gi.shape == oi.shape and np.array_equal(gi, oi))
        parity = {"checked": "collapsed mosaic after the first %d frames (one m2d_feed_batch, default group size) vs the CPU oracle" % ns,
                  "identical": same, "mosaic_px": int(gi.shape[0] * gi.shape[1]), "sha256_gpu": sha(gi), "sha256_oracle": sha(oi),
                  "differing_bytes": None if same or gi.shape != oi.shape else int((gi != oi).sum())}
        mp.close()

    m.close()
    del dev
    m2d.free_pinned(host_ptr)
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms_step, "gpu_launches": int(launches), "frames_fused": int(fused), "levels": levels,
            "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "config": config_of(mode, n, w, h, seed)}


def stream_latency(args, m2d, typ, local_rank, ns):
    """cfg5: one synchronous feed() per frame through the host-buffer API (H2D included), 1920x1080, pose jitter."""
    _, _, w, h, seed, jitter = CFG5
    s5 = synth.Sequence(ns, w, h, seed=seed, jitter=jitter)
    hf, hfp = m2d.pinned_empty((h, w, 3))
    ms5 = m2d.Map2D.create(typ, thread=False, device=local_rank)
    assert ms5.prepare(s5.plane, s5.camera, s5.prepare_poses)
    lat = []
    for k in range(ns):
        hf[:] = s5.frame(k)
        t0 = time.perf_counter()
        ms5.feed(hf, s5.poses[k])
        ms5.sync()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat[10:])  # first frames allocate staging buffers
    ms5.close()
    m2d.free_pinned(hfp)
    return {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "max_ms": float(lat.max()),
            "mean_ms": float(lat.mean()), "fps_sustained": float(1e3 / lat.mean()), "frames": int(len(lat))}


def device_frames(torch, seq, lo, hi, device):
    """Frames [lo, hi) of a synthetic sequence built ON the GPU: the same periodic-texture crops Sequence.frame() makes
    (an integer gather, bit-identical), without the host round trip (12 Mpx frames: 36 MB each)."""
    tex = torch.from_numpy(seq.texture).to(device)
    out = torch.empty((hi - lo, seq.h, seq.w, 3), dtype=torch.uint8, device=device)
    ar_h, ar_w = np.arange(seq.h), np.arange(seq.w)
    for k in range(lo, hi):
        r0, c0 = seq.frame_origin(k)
        rows = torch.from_numpy((r0 - ar_h) % tex.shape[0]).to(device)
        cols = torch.from_numpy((c0 + ar_w) % tex.shape[1]).to(device)
        out[k - lo] = tex[rows[:, None], cols[None, :]]
    return out


def measure_cfg3(args, m2d, torch, local_rank, stream):
    """BASELINE configs[2] on ONE GPU: the 1000-frame 4000x3000 multi-band survey, frames resident in HBM (36 GB)."""
    mode, n, w, h, seed, _ = CFG3
    n = args.cfg3_frames
    seq = synth.Sequence(n, w, h, seed=seed)
    dev = device_frames(torch, seq, 0, n, torch.device("cuda", local_rank))
    m = m2d.Map2D.create(3, thread=False, device=local_rank, batch_frames=args.batch)
    m.set_stream(stream.cuda_stream)
    assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)

    def step():
        m.reset()
        return m.feed_batch(dev.data_ptr(), n, w * h * 3, w, h, w * 3, seq.poses, True)

    for _ in range(2):
        res = step()
    m.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = max(2, min(args.steps, 5))
    ev0.record(stream)
    for _ in range(reps):
        res = step()
    ev1.record(stream)
    m.sync()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    fused = int((res == 0).sum())
    tiles = m.tile_count()
    # save half: collapse + D2H of the whole mosaic (the first call also allocates the collapse buffers: ~12 GB here)
    t0 = time.perf_counter()
    img, _ = m.get_image()
    save_cold_ms = (time.perf_counter() - t0) * 1e3
    out_pinned, out_ptr = m2d.pinned_empty((img.nbytes,))
    t0 = time.perf_counter()
    m.get_image(out=out_pinned)
    save_ms = (time.perf_counter() - t0) * 1e3
    m2d.free_pinned(out_ptr)
    out = {"workload": workload_name(mode, n, w, h, seed), "scaling": "strong", "n_gpus": 1, "value": fused * w * h / (ms * 1e-3) / 1e6,
           "unit": UNIT, "ms_per_step": ms, "steps": reps, "frames_fused": fused, "tiles": tiles, "tile_state_gb": tiles * m.tile_bytes() / 1e9,
           "save_ms": save_ms, "save_first_call_ms": save_cold_ms, "mosaic": [int(img.shape[1]), int(img.shape[0])], "mosaic_sha256": sha(img)}
    del img
    m.close()
    del dev
    torch.cuda.empty_cache()
    return out


def measure_render(args, m2d, torch, local_rank, stream, n=20):
    """Map2DRender (type 4), ONE batch of n 1280x720 frames (the reference renders its prepare-frames, 10-20 of
    them, as one batch) through m2d_render_frames: warp into per-frame boxes, sub-image pyramids, canvas blend, restore."""
    _, _, w, h, seed, _ = CFG2
    seq = synth.Sequence(n, w, h, seed=seed, jitter=True)
    host, host_ptr = m2d.pinned_empty((n, h, w, 3))
    for k in range(n):
        host[k] = seq.frame(k)
    dev = torch.from_numpy(host).cuda()
    out = {"workload": "Map2DRender batch: %d synthetic %dx%d frames, pose jitter, selection blend (what the reference executes), auto band count" % (n, w, h),
           "unit": UNIT}
    m = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=False, device=local_rank)
    m.set_stream(stream.cuda_stream)
    assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)

    def step():
        return m.render_frames(dev.data_ptr(), seq.poses, on_device=True, w=w, h=h)

    for _ in range(3):
        rc, res = step()
    m.sync()
    l0 = m.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = max(3, min(args.steps, 10))
    ev0.record(stream)
    for _ in range(reps):
        step()
    ev1.record(stream)
    m.sync()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    fused = int((res == 0).sum())
    r16, mask, nb, org = m.render_get()
    out.update({"value": fused * w * h / (ms * 1e-3) / 1e6, "ms_per_step": ms, "steps": reps, "frames_blended": fused, "bands": nb,
                "canvas": [int(r16.shape[1]), int(r16.shape[0])], "gpu_launches": int((m.launch_count() - l0) // reps)})
    # e2e: host frames in, 8-bit canvas out
    t0 = time.perf_counter()
    for _ in range(3):
        m.render_frames(host, seq.poses)
        img, _ = m.get_image()
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    out["e2e"] = {"value": fused * w * h / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(host.nbytes),
                  "d2h_bytes_per_step": int(img.nbytes)}
    if not args.no_cpu:
        from oracle import oracle as O
        threads = os.cpu_count() or 1
        O.set_threads(threads)
        o = O.OracleMap2D(O.TYPE_RENDER)
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        t0 = time.perf_counter()
        o.render_frames(host, seq.poses)
        dt = time.perf_counter() - t0
        O.set_threads(1)
        ref16, refmask, refnb, reforg = o.render_get()
        same = bool(refnb == nb and reforg == org and np.array_equal(ref16, r16) and np.array_equal(refmask, mask))
        out["cpu_baseline"] = {"value": fused * w * h / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": "the whole batch, oracle/render_oracle.inl (warps and pyramids OpenMP, blend loop serial; %.1f s)" % dt}
        out["parity"] = {"checked": "the blender's CV_16SC3 result + mask of the whole batch vs the CPU oracle", "identical": same,
                         "sha256_gpu": sha(r16), "sha256_oracle": sha(ref16)}
    m.close()
    del dev
    m2d.free_pinned(host_ptr)
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default="multiband", choices=["multiband", "weighted"], help="which mode is the headline")
    ap.add_argument("--frames", type=int, default=0, help="frames of the headline workload (default: the BASELINE config's)")
    ap.add_argument("--size", default="", help="WxH of the headline workload's frames (default 1280x720)")
    ap.add_argument("--ref-frames", type=int, default=0, help="reference arm: frames per step (default: ~460 Mpx worth, 500 at 720p)")
    ap.add_argument("--cpu-frames", type=int, default=48)
    ap.add_argument("--batch", type=int, default=0, help="m2d_config.batch_frames (0 = library default)")
    ap.add_argument("--stream-latency", type=int, default=CFG5[1], help="frames of the cfg5 latency run (0 = skip)")
    ap.add_argument("--cfg3-frames", type=int, default=CFG3[1], help="frames of the cfg3 run (0 = skip)")
    ap.add_argument("--cfg4-frames", type=int, default=-1, help="frames of the cfg4 run (N>1 only; default: 5000 at 8 GPUs, else skipped; 0 = skip)")
    ap.add_argument("--only", action="store_true", help="headline mode only: no second mode, no cfg3, no cfg5")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1 headline. weak: N x the survey, --frames per GPU, strips of tiles, halo frames by P2P; strong: the same frames cut into N shards")
    ap.add_argument("--halo", default="peer", choices=["peer", "copy"],
                    help="N>1: how halo frames reach a rank. peer: sampled in place from the neighbour's HBM over NVLink (CUDA IPC); copy: NCCL P2P copies every step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--zero-copy", action="store_true", help="e2e experiment: pass the pinned host frames as device pointers (sampled in place over PCIe)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import pi_slam_fusion_b200.map2d as m2d
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        from pi_slam_fusion_b200 import sharded
        sharded.bench_main(args, rank, world, local_rank)
        return

    stream = torch.cuda.Stream()
    head_cfg, other_cfg = (CFG2, CFG1) if args.mode == "multiband" else (CFG1, CFG2)
    mode, n, w, h, seed, _ = head_cfg
    n = args.frames or n
    if args.size:
        w, h = (int(v) for v in args.size.lower().split("x"))
    head = measure_mode(args, m2d, torch, mode, n, w, h, seed, local_rank, stream)
    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16" if mode == "multiband" else "u8", "data": "synthetic", "config": head["config"],
            "run": {"frames_fused": head["frames_fused"], "batch_frames": args.batch, "parallelism": "1 GPU"},
            "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "cpu_baseline": head["cpu_baseline"], "parity": head["parity"]}
    if not args.only:
        omode, on, ow, oh, oseed, _ = other_cfg
        o = measure_mode(args, m2d, torch, omode, on, ow, oh, oseed, local_rank, stream, want_clocks=True)
        line[omode] = {k: o[k] for k in ("value", "ms_per_step", "gpu_launches", "frames_fused", "clocks", "e2e", "roofline", "cpu_baseline", "parity", "config")}
        line[omode]["unit"] = UNIT
        line[omode]["dtype"] = "s16" if omode == "multiband" else "u8"
        if args.stream_latency > 0:
            sl = {"workload": "cfg5: %d synchronous feed() calls, 1920x1080, pose jitter, 80 %% overlap, host->device copy included" % args.stream_latency}
            for name, typ in (("multiband", 3), ("weighted", 1)):
                sl[name] = stream_latency(args, m2d, typ, local_rank, args.stream_latency)
            line["stream_latency"] = sl
        if args.cfg3_frames > 0:
            line["cfg3"] = measure_cfg3(args, m2d, torch, local_rank, stream)
        try:
            line["render"] = measure_render(args, m2d, torch, local_rank, stream)
        except Exception as e:   # a sub-run must not cost the headline line
            line["render"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
