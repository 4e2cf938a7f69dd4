"""the on-disk dataset format of Map2DFusion (config.cfg / trajectory.txt / rgb/) and the headless
replay driver.  CPU: round trip of the format and a replay through the oracle; GPU: the same dataset through the
CUDA path equals the oracle, and the CLI writes a PNG."""
import os
import subprocess
import sys

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as O  # noqa: E402
import pi_slam_fusion_b200.synth as synth  # noqa: E402
from pi_slam_fusion_b200 import replay  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_dataset(tmp_path, n=8):
    seq = synth.Sequence(n, 320, 180, seed=31, jitter=True, fpl=4, prepare_frames=4)
    plane = np.array([1.5, -2.0, 0.25, 0.0, 0.0, np.sin(0.05), np.cos(0.05)])  # a non-identity ground plane
    d = str(tmp_path / "ds")
    replay.write_dataset(d, plane, seq.camera, seq.poses, seq.frame, gps_origin=[108.888931, 34.257287, 400.0])
    return d, seq, plane


def test_format_round_trip(tmp_path):
    d, seq, plane = make_dataset(tmp_path)
    ds = replay.Dataset(d)
    assert len(ds) == seq.n
    assert np.array_equal(ds.poses, seq.poses)                       # %.17g round-trips doubles
    assert np.allclose(ds.plane, plane, rtol=0, atol=1e-9) and np.array_equal(ds.camera, seq.camera)
    assert ds.gps_origin == [108.888931, 34.257287, 400.0]
    assert np.array_equal(ds.image(3), seq.frame(3))                 # lossless png
    assert ds.trajectory_length() > 0
    cfg = open(os.path.join(d, "config.cfg")).read()
    assert "Camera.Paraments=[320 180 288 288 160 90]" in cfg and cfg.startswith("Plane=")


def test_reference_style_config_is_parsed(tmp_path):
    p = tmp_path / "config.cfg"
    p.write_text("Plane=0 0 0 0 0 0 1\nCamera.CameraType=PinHole\nCamera.Paraments=[1920 1080 1184.5 1183.9 978.4 533.8]\n"
                 "TrajectoryFile=$(Svar.ParsingPath)/trajectory.txt\nGPS.Origin=108.9 34.2 0\n// comment\nPrepareFrameNum?=10\n")
    (tmp_path / "trajectory.txt").write_text("1438158112.560000 1 2 3 0 0 0 1\n")
    ds = replay.Dataset(str(tmp_path))
    assert list(ds.camera) == [1920, 1080, 1184.5, 1183.9, 978.4, 533.8] and ds.stamps == ["1438158112.560000"]
    assert ds.cfg["PrepareFrameNum"] == "10"
    (tmp_path / "config.cfg").write_text("Plane=0 0 0 0 0 0 1\nCamera.Paraments=[1 2 3]\n")
    with pytest.raises(ValueError):
        replay.Dataset(str(tmp_path))


def test_replay_through_oracle_equals_direct_feed(tmp_path):
    d, seq, plane = make_dataset(tmp_path)
    ds = replay.Dataset(d)
    for typ in (1, 3):
        a = O.OracleMap2D.create(typ)
        acc = replay.replay(ds, a, prepare_frames=4)
        b = O.OracleMap2D.create(typ)
        assert b.prepare(plane, seq.camera, seq.poses[:4])
        assert acc == [b.feed(seq.frame(k), seq.poses[k]) for k in range(seq.n)] and all(acc)
        assert np.array_equal(a.get_image()[0], b.get_image()[0])
        c = O.OracleMap2D.create(typ)     # frame-by-frame feed() and grouped feed_batch() replay agree
        assert replay.replay(ds, c, prepare_frames=4, batch=1) == acc == replay.replay(ds, O.OracleMap2D.create(typ), 4, batch=3)
        assert np.array_equal(a.get_image()[0], c.get_image()[0])


@pytest.mark.gpu
@pytest.mark.parametrize("typ", [1, 3])
def test_replay_on_gpu_equals_oracle(tmp_path, typ):
    import pi_slam_fusion_b200.map2d as m2d
    d, seq, plane = make_dataset(tmp_path)
    ds = replay.Dataset(d)
    g = m2d.Map2D.create(typ, thread=False)
    o = O.OracleMap2D.create(typ)
    assert replay.replay(ds, g, 4) == replay.replay(ds, o, 4)
    gi, oi = g.get_image(), o.get_image()
    assert gi[1] == oi[1] and np.array_equal(gi[0], oi[0])
    g.close()


@pytest.mark.gpu
def test_replay_cli(tmp_path):
    d, seq, plane = make_dataset(tmp_path)
    out = str(tmp_path / "result.png")
    r = subprocess.run([sys.executable, "-m", "pi_slam_fusion_b200.replay", d, "--type", "3", "--out", out, "--prepare", "4"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "fused 8" in r.stdout and os.path.getsize(out) > 1000
    img = cv2.imread(out, cv2.IMREAD_UNCHANGED)
    assert img.shape[0] % 256 == 0 and img.shape[2] == 3
