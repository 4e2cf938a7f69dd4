"""One process for the round's ncu captures: (1) one 250-frame group of cfg2 with frames resident in HBM, (2) the same group with
frames in pinned HOST memory (pull mode: mbs_mark / mbs_pull), (3) one Map2DRender batch of 20 frames.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python scripts/prof_all.py
    ncu --set full --clock-control none -k regex:'...' -o rep python scripts/prof_all.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pi_slam_fusion_b200.map2d as m2d  # noqa: E402
import pi_slam_fusion_b200.synth as synth  # noqa: E402

N, W, H = 250, 1280, 720
seq = synth.Sequence(500, W, H, seed=2)
host, host_ptr = m2d.pinned_empty((N, H, W, 3))
for k in range(N):
    host[k] = seq.frame(k)
dev = torch.from_numpy(host).cuda()
m = m2d.Map2D.create(3, thread=False)
assert m.prepare(seq.plane, seq.camera, seq.prepare_poses)
res = m.feed_batch(dev.data_ptr(), N, W * H * 3, W, H, W * 3, seq.poses[:N], True)
m.sync()
print("device-resident: fed", int((res == 0).sum()), "launches", m.launch_count())
m.reset()
l0 = m.launch_count()
res = m.feed_batch(host_ptr, N, W * H * 3, W, H, W * 3, seq.poses[:N], False)
m.sync()
print("pinned host (pull mode): fed", int((res == 0).sum()), "launches", m.launch_count() - l0)
m.close()
del dev
rs = synth.Sequence(20, W, H, seed=2, jitter=True)
rdev = torch.from_numpy(rs.frames()).cuda()
r = m2d.Map2D.create(m2d.Map2D.TypeRender, thread=False)
assert r.prepare(rs.plane, rs.camera, rs.prepare_poses)
rc, rres = r.render_frames(rdev.data_ptr(), rs.poses, on_device=True, w=W, h=H)
r.sync()
print("render: blended", int((rres == 0).sum()), "bands", r.render_get()[2], "launches", r.launch_count())
