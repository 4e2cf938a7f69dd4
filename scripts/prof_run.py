"""Small driver for ncu captures: feed N frames of the bench workload through the C-ABI (device-resident frames).

    python scripts/prof_run.py --mode multiband --frames 24 [--batch]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pi_slam_fusion_b200.map2d as m2d  # noqa: E402
import pi_slam_fusion_b200.synth as synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="multiband")
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--w", type=int, default=1280)
ap.add_argument("--h", type=int, default=720)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--pinned", action="store_true", help="frames in pinned HOST memory (pull mode / in-place sampling) instead of HBM")
a = ap.parse_args()
if a.mode == "render":   # Map2DRender (type 4): one batch through m2d_render_frames
    seq = synth.Sequence(a.frames, a.w, a.h, seed=2, jitter=True)
    dev = torch.from_numpy(seq.frames()).cuda()
    m = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=False)
    assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)
    for _ in range(a.reps):
        rc, res = m.render_frames(dev.data_ptr(), seq.poses, on_device=True, w=a.w, h=a.h)
        m.sync()
    print("rendered", int((res == 0).sum()), "frames,", m.render_get()[2], "bands, launches", m.launch_count())
    sys.exit(0)
typ = 3 if a.mode == "multiband" else 1
seq = synth.Sequence(500, a.w, a.h, seed=2)
frames = np.stack([seq.frame(k) for k in range(a.frames)])
if a.pinned:
    host, host_ptr = m2d.pinned_empty(frames.shape)
    host[:] = frames
else:
    dev = torch.from_numpy(frames).cuda()
m = m2d.Map2D.create(typ, thread=False)
assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)
for _ in range(a.reps):
    m.reset()
    res = m.feed_batch(host_ptr if a.pinned else dev.data_ptr(), a.frames, a.w * a.h * 3, a.w, a.h, a.w * 3, seq.poses[:a.frames], not a.pinned)
    m.sync()
print("fed", int((res == 0).sum()), "frames, launches", m.launch_count())
