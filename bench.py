#!/usr/bin/env python
"""bench.py — fused input Mpix/s of Map2D::feed() (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode multiband|weighted]

A "step" is one pass of the hot path over the whole synthetic survey of BASELINE.json configs[1]
(MultiBandMap2DCPU, 5 bands, 500 nadir 1280x720 frames, seed 2): reset the map, feed every frame.
  value     whole-job Mpix/s with the frames already resident in HBM (feed_batch on device pointers)
  e2e       the same through the public host-buffer API: frames in pinned host memory, H2D copies inside the
            timed region, plus the collapsed mosaic (Map2D::save in memory) read back to the host every step
  roofline  dominant kernel: stage-algorithmic bytes per launch / CUDA-event launch duration (measured live via
            m2d_profile) against MEASURED_PEAKS.json; "path" = SURVEY §8(d) whole-frame algorithmic bytes / step time
  cpu_baseline  the CPU oracle (a port of the reference recipe), 1 thread, on a bounded sample of the same frames
`--impl reference` times the oracle with all host threads instead (the reference cannot be compiled here, see
DESIGN.md); under torchrun only rank 0 works.
"""
import argparse
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

W, H, NFRAMES, SEED = 1280, 720, 500, 2
METRIC = "fused input Mpix/s"
UNIT = "Mpix/s"


def workload_name(mode, n):
    if mode == "multiband":
        return "cfg2: MultiBandMap2DCPU 5-band Laplacian blend, %d synthetic %dx%d BGR nadir frames (serpentine, seed %d)" % (n, W, H, SEED)
    return "cfg1-shaped: Map2DCPU weighted fusion, %d synthetic %dx%d BGR nadir frames (serpentine, seed %d)" % (n, W, H, SEED)


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
            time.sleep(0.02)

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


def stage_bytes(mode, stats, levels):
    """Per kernel class, per STEP: bytes the stage must move given ITS inputs/outputs (scratch pyramid counted at the
    reference's element sizes: 6 B int16x3 + 4 B f32 per px).  Divided by the class's launch count -> per launch."""
    out = {}
    if mode == "weighted":
        out["weighted_fuse"] = algorithmic_bytes(mode, stats, levels)
        return out
    D = [stats["region_px"][l] for l in range(levels)]
    out["mb_warp"] = 3 * stats["input_px"] + 10 * D[0]
    # pyrDown bytes are reported for both pyramid kernel classes together under "mb_pyrdown" (which levels go to the
    # tail kernel depends on the frame size); the tiny tail gets the deepest level's share
    out["mb_pyrdown"] = sum(10 * D[l] + 10 * D[l + 1] for l in range(levels - 2))
    out["mb_pyrtail"] = sum(10 * D[l] + 10 * D[l + 1] for l in range(max(levels - 2, 0), levels - 1))
    sel = dec = lap = 0.0
    for l in range(levels):
        nonfresh = stats["region_px"][l] - stats["fresh_px"][l]
        written = stats["fresh_px"][l] + stats["win_px"][l]
        sel += 4 * D[l] + 4 * nonfresh + written * (6 + 10 + (6 / 4 if l + 1 < levels else 0))
        dec += 4 * D[l] + 4 * nonfresh + 4 * written           # every frame's weight, the tile weight, the new weight
        lap += written * (6 + 6 + (6 / 4 if l + 1 < levels else 0))  # winner's G_l, its share of G_{l+1}, the Laplacian out
    out["mb_select"] = sel
    # weights-first pipeline (default): weights dense (4 B f32 per px), image only where a winner needs it.  need_px[l] =
    # px of the frames' level l that had to be computed (counted exactly by the library with collect_stats); the image
    # stages are charged the reference's 6 B int16x3 per needed px, the source at 3 B per needed level-0 px.
    N = stats.get("need_px", [0] * levels)
    out["mbw_warp"] = 4 * D[0]
    out["mbw_pyramid"] = sum(4 * D[l] + 4 * D[l + 1] for l in range(levels - 1))
    out["mbs_decide"] = dec
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


def run_reference(args, rank):
    """Reference arm: the CPU oracle (port of the reference recipe), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    mode = args.mode
    typ = 3 if mode == "multiband" else 1
    threads = os.cpu_count() or 1
    O.set_threads(threads)
    seq = synth.Sequence(NFRAMES, W, H, seed=SEED)
    sample = args.ref_frames
    frames = [seq.frame(k) for k in range(sample)]

    def step():
        o = O.OracleMap2D.create(typ)
        assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
        for k in range(sample):
            o.feed(frames[k], seq.poses[k])

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample * W * H / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "s16" if mode == "multiband" else "u8", "data": "synthetic",
            "config": {"workload": workload_name(mode, NFRAMES), "mode": mode},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "first %d frames of the %d-frame workload per step (oracle/map2d_oracle.cpp, OpenMP)" % (sample, NFRAMES)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default="multiband", choices=["multiband", "weighted"])
    ap.add_argument("--frames", type=int, default=NFRAMES)
    ap.add_argument("--size", default="", help="WxH of the synthetic frames (default 1280x720 = BASELINE configs[1]); e.g. 4000x3000 for cfg3-sized frames")
    ap.add_argument("--ref-frames", type=int, default=24)
    ap.add_argument("--cpu-frames", type=int, default=48)
    ap.add_argument("--batch", type=int, default=0, help="m2d_config.batch_frames (0 = library default)")
    ap.add_argument("--stream-latency", type=int, default=0,
                    help="also measure synchronous per-frame feed() latency (BASELINE cfg5: 1920x1080, pose jitter, H2D included) over N frames")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1 only. weak: N x the survey, --frames per GPU, strips of tiles, halo frames by P2P; strong: the same --frames cut into N shards")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.size:
        global W, H
        W, H = (int(v) for v in args.size.lower().split("x"))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
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

    mode = args.mode
    typ = 3 if mode == "multiband" else 1
    n = args.frames
    seq = synth.Sequence(n, W, H, seed=SEED)
    frame_bytes = W * H * 3
    host, host_ptr = m2d.pinned_empty((n, H, W, 3))
    for k in range(n):
        host[k] = seq.frame(k)
    dev = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    dev.copy_(torch.from_numpy(host))
    torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    m = m2d.Map2D.create(typ, thread=False, device=local_rank, batch_frames=args.batch)
    m.set_stream(stream.cuda_stream)
    assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)

    def step_device():
        m.reset()
        res = m.feed_batch(dev.data_ptr(), n, frame_bytes, W, H, W * 3, seq.poses, True)
        return res

    # ---- exact algorithmic bytes from one counted pass (separate handle, outside every timed region)
    mc = m2d.Map2D.create(typ, thread=False, device=local_rank, collect_stats=1, batch_frames=args.batch)
    assert mc.prepare(seq.plane, seq.camera, seq.prepare_poses)
    mc.feed_batch(dev.data_ptr(), n, frame_bytes, W, H, W * 3, seq.poses, True)
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
    clocks = sampler.result()
    launches = (m.launch_count() - l0) // args.steps
    value = fused * W * H / (ms_step * 1e-3) / 1e6

    # ---- live per-kernel timing (CUDA events around every launch on the launching stream)
    m.reset()
    m.profile(True)
    m.feed_batch(dev.data_ptr(), n, frame_bytes, W, H, W * 3, seq.poses, True)
    kt = m.kernel_times()
    m.profile(False)
    peak, peak_src = peaks()
    sb = stage_bytes(mode, stats, levels)
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
    roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": (per_kernel[dom]["achieved_gbs"] or 0) / peak, "traffic": ncu_traffic(dom), "peak_source": peak_src,
                "kernels": per_kernel,
                "path": {"bytes_per_step": path_bytes, "bytes_per_input_px": path_bytes / (fused * W * H),
                         "achieved": path_ach, "frac": path_ach / peak}}

    # ---- e2e: host buffers through the public API, H2D inside, mosaic read back
    e2e = None
    if not args.no_e2e:
        me = m2d.Map2D.create(typ, thread=False, device=local_rank, batch_frames=args.batch)
        me.set_stream(stream.cuda_stream)
        assert me.prepare(seq.plane, seq.camera, seq.prepare_poses)
        me.feed_batch(host_ptr, n, frame_bytes, W, H, W * 3, seq.poses, False)
        img, _ = me.get_image()
        out_bytes = img.nbytes
        del img
        out_pinned, out_ptr = m2d.pinned_empty((out_bytes,))

        def step_e2e():
            me.reset()
            me.feed_batch(host_ptr, n, frame_bytes, W, H, W * 3, seq.poses, False)
            return me.get_image(out=out_pinned)

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
        e2e = {"value": fused * W * H / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": n * frame_bytes,
               "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_e2e,
               "what": "m2d_feed_batch(host pinned frames) + m2d_get_image (collapse + D2H of the mosaic)"}
        me.close()
        m2d.free_pinned(out_ptr)

    # ---- CPU baseline: the oracle, one thread, bounded sample
    cpu = None
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
        cpu = {"value": ns * W * H / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "first %d frames of the workload, oracle/map2d_oracle.cpp single thread (%.1f s)" % (ns, dt),
               "host_cores": os.cpu_count()}

    # ---- optional: streaming latency (cfg5), one synchronous feed() per frame through the host-buffer API
    stream_lat = None
    if args.stream_latency > 0:
        ns = args.stream_latency
        s5 = synth.Sequence(ns, 1920, 1080, seed=5, jitter=True)
        hf, hfp = m2d.pinned_empty((1080, 1920, 3))
        ms5 = m2d.Map2D.create(typ, thread=False, device=local_rank)
        assert ms5.prepare(s5.plane, s5.camera, s5.prepare_poses)
        lat = []
        for k in range(ns):
            hf[:] = s5.frame(k)
            t0 = time.perf_counter()
            ms5.feed(hf, s5.poses[k])
            ms5.sync()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat[10:])  # first frames allocate pool slabs and staging buffers
        stream_lat = {"workload": "cfg5: %d synchronous feed() calls, 1920x1080, pose jitter, host->device copy included" % ns,
                      "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean()),
                      "fps_sustained": float(1e3 / lat.mean())}
        ms5.close()
        m2d.free_pinned(hfp)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "s16" if mode == "multiband" else "u8", "data": "synthetic",
            "config": {"workload": workload_name(mode, n), "mode": mode, "frames": n, "frames_fused": fused,
                       "frame": [W, H], "bands": levels - 1, "l2": "inputs %.2f GB per step > 126 MB L2 (no flush needed)" % (n * frame_bytes / 1e9),
                       "parallelism": "1 GPU", "batch_frames": args.batch},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    if stream_lat:
        line["stream_latency"] = stream_lat
    print(json.dumps(line))
    m.close()
    m2d.free_pinned(host_ptr)


if __name__ == "__main__":
    main()
