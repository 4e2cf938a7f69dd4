"""Headless replay driver + on-disk dataset format of Map2DFusion.

Replaces `TestSystem::testMap2D` (reference `Map2DFusion/Map2DFusion.cpp:153-248`, commented original reader
`:116-137`) without Qt: a dataset directory as written by `MapHash::saveMap2DFusion`
(`GSLAM-DIYSLAM/src/zhaoyong/MapHash.cpp:655-760`):

    config.cfg        Plane=x y z qx qy qz qw
                      Camera.CameraType=PinHole
                      Camera.Paraments=[w h fx fy cx cy]
                      [GPS.Origin=lon lat alt]  [TrajectoryFile=...]
    trajectory.txt    one line per frame:  <timestamp> x y z qx qy qz qw      (pose = camera-to-world, SE3.h:105-117)
    rgb/<timestamp>.jpg   (any extension cv2.imread understands; tests use lossless .png)

    python -m pi_slam_fusion_b200.replay DATASET --type 3 --out result.png [--prepare 10] [--batch 32]

The driver mirrors the reference loop: create(type) -> prepare(plane, camera, first PrepareFrameNum frames) ->
feed every frame -> save(Map.File2Save).  (thread=false convention: the prepare-frames are fed explicitly.)
"""
import argparse
import os
import re
import sys

import numpy as np

_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")


def _numbers(s):
    return [float(x) for x in _NUM.findall(s)]


def parse_config(path):
    """The subset of the svar .cfg mini-language the dataset writer emits: `key=value` lines, `//` and `#` comments."""
    cfg = {}
    with open(path) as f:
        for line in f:
            line = line.split("//")[0].split("#")[0].strip()
            if "=" not in line:
                continue
            k, v = line.split("=", 1)
            cfg[k.strip().rstrip("?")] = v.strip()
    return cfg


class Dataset:
    def __init__(self, path):
        self.path = path
        cfg = parse_config(os.path.join(path, "config.cfg"))
        self.cfg = cfg
        self.plane = np.array(_numbers(cfg["Plane"]), np.float64) if "Plane" in cfg else np.array([0, 0, 0, 0, 0, 0, 1.0])
        cam = _numbers(cfg.get("Camera.Paraments", ""))
        if len(cam) != 6:
            raise ValueError("Invalid camera parameters!")  # Map2DFusion.cpp:196-200
        if self.plane.shape != (7,):
            raise ValueError("Plane must be 7 numbers (x y z qx qy qz qw)")
        self.camera = np.array(cam, np.float64)
        self.gps_origin = _numbers(cfg["GPS.Origin"]) if "GPS.Origin" in cfg else None
        traj = cfg.get("TrajectoryFile", "").replace("$(Svar.ParsingPath)", path) or os.path.join(path, "trajectory.txt")
        if not os.path.exists(traj):
            traj = os.path.join(path, "trajectory.txt")
        self.stamps, poses = [], []
        with open(traj) as f:
            for line in f:
                tok = line.split()
                if len(tok) < 8:
                    continue
                self.stamps.append(tok[0])
                poses.append([float(t) for t in tok[1:8]])
        self.poses = np.array(poses, np.float64).reshape(-1, 7)

    def __len__(self):
        return len(self.stamps)

    def image_path(self, k):
        base = os.path.join(self.path, "rgb", self.stamps[k])
        for ext in (".jpg", ".png", ".jpeg", ".bmp", ".ppm"):
            if os.path.exists(base + ext):
                return base + ext
        return base + ".jpg"

    def image(self, k):
        import cv2
        img = cv2.imread(self.image_path(k), cv2.IMREAD_COLOR)  # BGR, like the reference's cv::imread
        if img is None:
            raise IOError("cannot read " + self.image_path(k))
        return img

    def trajectory_length(self):
        """TrajectoryLengthCalculator (Map2DFusion.cpp:14-35): summed distance between consecutive camera centres."""
        d = np.diff(self.poses[:, :3], axis=0)
        return float(np.sqrt((d * d).sum(1)).sum())


def write_dataset(path, plane, camera, poses, frames, ext=".png", stamps=None, gps_origin=None):
    """Writes the same layout as MapHash::saveMap2DFusion (timestamps with 6 decimals, precision 10)."""
    import cv2
    os.makedirs(os.path.join(path, "rgb"), exist_ok=True)
    poses = np.asarray(poses, np.float64).reshape(-1, 7)
    stamps = stamps or ["%.6f" % (1000.0 + 0.1 * k) for k in range(len(poses))]
    with open(os.path.join(path, "config.cfg"), "w") as f:
        f.write("Plane=" + " ".join("%.10g" % v for v in plane) + "\n")
        f.write("Camera.CameraType=PinHole\n")
        f.write("Camera.Paraments=[" + " ".join("%.10g" % v for v in camera) + "]\n")
        f.write("TrajectoryFile=$(Svar.ParsingPath)/trajectory.txt\n")
        if gps_origin is not None:
            f.write("GPS.Origin=" + " ".join("%.10g" % v for v in gps_origin) + "\n")
    with open(os.path.join(path, "trajectory.txt"), "w") as f:
        for s, p in zip(stamps, poses):
            f.write(s + " " + " ".join("%.17g" % v for v in p) + "\n")
    for k, s in enumerate(stamps):
        img = frames(k) if callable(frames) else frames[k]
        if not cv2.imwrite(os.path.join(path, "rgb", s + ext), img):
            raise IOError("cannot write image %d" % k)
    return stamps


def replay(dataset, map2d, prepare_frames=10, batch=32, on_frame=None):
    """create()'d map object in, fused map out.  Returns the list of per-frame accept flags.

    `map2d` needs prepare(plane, camera, poses) and feed(img, pose); the product passes map2d.Map2D."""
    n = len(dataset)
    if n == 0:
        raise ValueError("empty trajectory")
    npre = min(max(prepare_frames, 1), n)
    if not map2d.prepare(dataset.plane, dataset.camera, dataset.poses[:npre]):
        raise RuntimeError("prepare() rejected the dataset (camera heights straddle the plane?)")
    accepted = []
    batch = max(int(batch), 1)
    if batch > 1 and hasattr(map2d, "feed_batch"):
        # grouped launches: `batch` decoded frames at a time through m2d_feed_batch (host buffers); same results as
        # frame-by-frame feed() calls, each map tile touched once per group instead of once per frame
        for k0 in range(0, n, batch):
            imgs = []
            for k in range(k0, min(k0 + batch, n)):
                img = dataset.image(k)
                if img is None or img.shape[:2] != (int(dataset.camera[1]), int(dataset.camera[0])) or img.dtype != np.uint8 or img.ndim != 3:
                    img = None   # renderFrame would reject it (Map2DCPU.cpp:158-162)
                imgs.append(img)
            good = [i for i, im in enumerate(imgs) if im is not None]
            flags = [False] * len(imgs)
            if good:
                stack = np.ascontiguousarray(np.stack([imgs[i] for i in good]))
                h, w = stack.shape[1:3]
                res = map2d.feed_batch(stack.ctypes.data, len(good), w * h * 3, w, h, w * 3, dataset.poses[[k0 + i for i in good]], False)
                if hasattr(map2d, "sync"):
                    map2d.sync()   # `stack` is read until here
                for i, r in zip(good, res):
                    flags[i] = (int(r) == 0)
            for i, ok in enumerate(flags):
                accepted.append(ok)
                if on_frame:
                    on_frame(k0 + i, ok)
        return accepted
    for k in range(n):
        ok = map2d.feed(dataset.image(k), dataset.poses[k])
        accepted.append(bool(ok))
        if on_frame:
            on_frame(k, ok)
    if hasattr(map2d, "sync"):
        map2d.sync()
    return accepted


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("dataset")
    ap.add_argument("--type", type=int, default=3, help="Map2D.Type: 1 weighted (Map2DCPU), 3 multi-band (MultiBandMap2DCPU)")
    ap.add_argument("--out", default="result.png", help="Map.File2Save")
    ap.add_argument("--prepare", type=int, default=10, help="PrepareFrameNum")
    ap.add_argument("--scale", type=float, default=1.0, help="Map2D.Scale")
    ap.add_argument("--resolution", type=float, default=0.0, help="Map2D.Resolution (multi-band)")
    ap.add_argument("--bands", type=int, default=5, help="MultiBandMap2DCPU.BandNumber")
    ap.add_argument("--weight-type", type=int, default=0, help="Map2D.WeightType")
    ap.add_argument("--background", type=int, default=0, help="Result.BackGroundColor")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--batch", type=int, default=32, help="frames handed to m2d_feed_batch at a time (1 = one feed() per frame)")
    a = ap.parse_args(argv)
    from . import map2d as m2d
    ds = Dataset(a.dataset)
    m = m2d.Map2D.create(a.type, thread=False, scale=a.scale, resolution=a.resolution, band_number=a.bands,
                         weight_type=a.weight_type, background=a.background, device=a.device)
    acc = replay(ds, m, a.prepare, batch=a.batch)
    ok = m.save(a.out)
    g = m.grid()
    print("frames %d fused %d  grid %dx%d tiles  lengthPixel %.6g  trajectory %.1f m  saved=%s -> %s"
          % (len(ds), sum(acc), g["w"], g["h"], g["length_pixel"], ds.trajectory_length(), ok, a.out))
    m.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
