"""Generates tests/golden/golden.json and kat.npz.  Run in the build container (needs cv2 4.13):

    python tests/golden/make_golden.py

Source of truth = the REAL OpenCV primitives through tests/cv2_reference.py (the reference itself cannot be
built here and ships no golden vectors).  Two families:
  "cv2":       produced by cv2 alone; the oracle must reproduce them (multi-band state in set_f32_mode(1)).
  "oracle249": produced by the oracle in its default OpenCV-2.4.9 float association (after the oracle has been
               checked against "cv2" in this same script); the CUDA path must reproduce them on the GPU box.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402

from oracle import oracle as O  # noqa: E402
import pi_slam_fusion_b200.synth as synth  # noqa: E402
from tests.cv2_reference import Cv2Map2D, warp_nearest_249  # noqa: E402

cv2.setNumThreads(1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


SEQS = {
    "small_jitter": dict(n=12, w=320, h=180, seed=7, jitter=True, noise=True, fpl=4, prepare_frames=6),
    "small_nadir": dict(n=10, w=256, h=144, seed=3, jitter=False, noise=False, fpl=5, prepare_frames=4),
}


def state_record(m, typ, levels):
    rec = {}
    g = m.grid() if hasattr(m, "grid") else {"w": m.w, "h": m.h}
    for ty in range(g["h"]):
        for tx in range(g["w"]):
            if typ == 1:
                t = m.get_tile(tx, ty)
                if t is not None:
                    rec["%d,%d" % (tx, ty)] = sha(t)
            else:
                t0 = m.get_tile(tx, ty, 0)
                if t0 is None:
                    continue
                rec["%d,%d" % (tx, ty)] = [[sha(m.get_tile(tx, ty, l)[0]), sha(m.get_tile(tx, ty, l)[1])] for l in range(levels)]
    return rec


def main():
    out = {"opencv": cv2.__version__, "sequences": SEQS, "cv2": {}, "oracle249": {}}
    kat = {}
    # ---- primitive KATs
    rng = np.random.default_rng(1234)
    src4 = np.array([[0, 0], [128, 0], [0, 72], [128, 72]], np.float32)
    dst4 = (src4 * 1.37 + np.array([[33.3, 41.7]]) + rng.normal(0, 3, (4, 2))).astype(np.float32)
    M = cv2.getPerspectiveTransform(src4, dst4)
    kat["H_src"], kat["H_dst"], kat["H"] = src4, dst4, M
    kat["H_inv"] = cv2.invert(M)[1]
    s8 = rng.integers(0, 256, (72, 128, 4), dtype=np.uint8)
    s16 = rng.integers(0, 256, (72, 128, 3)).astype(np.int16)
    sf = rng.random((72, 128), dtype=np.float32)
    kat["warp_src_u8c4"], kat["warp_src_s16c3"], kat["warp_src_f32"] = s8, s16, sf
    prim = {"warp_u8c4_256": sha(cv2.warpPerspective(s8, M, (256, 256), flags=cv2.INTER_LINEAR)),
            "warp_s16c3_reflect_256": sha(cv2.warpPerspective(s16, M, (256, 256), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)),
            "warp_f32_nearest_249_256": sha(warp_nearest_249(sf, M, (256, 256)))}
    for shape in ((16, 16), (5, 7), (64, 48)):
        a = rng.integers(-3000, 3000, shape + (3,)).astype(np.int16)
        f = rng.random(shape, dtype=np.float32)
        key = "%dx%d" % shape
        kat["pyr_s16_" + key], kat["pyr_f32_" + key] = a, f
        kat["pyrdown_s16_" + key], kat["pyrup_s16_" + key] = cv2.pyrDown(a), cv2.pyrUp(a)
        kat["pyrdown_f32_cv2_" + key] = cv2.pyrDown(f)
        O.set_f32_mode(0)
        kat["pyrdown_f32_249_" + key] = O.pyrdown_f32(f)
    out["cv2"]["primitives"] = prim
    # ---- sequences
    for name, kw in SEQS.items():
        seq = synth.Sequence(**kw)
        for typ in (1, 3):
            c = Cv2Map2D(typ)
            assert c.prepare(seq.plane, seq.camera, seq.prepare_poses)
            rects, accepted = [], []
            for k in range(seq.n):
                ok = c.feed(seq.frame(k), seq.poses[k])
                accepted.append(bool(ok))
                rects.append(list(c.last_rect) if ok else None)
            img, org = c.get_image()
            rec = {"grid": {"w": c.w, "h": c.h, "min": list(c.min), "max": list(c.max), "length_pixel": c.lp},
                   "accepted": accepted, "rects": rects, "tiles": state_record(c, typ, 6), "mosaic": sha(img),
                   "mosaic_origin": list(org), "mosaic_shape": list(img.shape)}
            out["cv2"]["%s/type%d" % (name, typ)] = rec
            # the oracle must reproduce cv2 (mode 1), then we record its default-mode (2.4.9) state
            O.set_f32_mode(1)
            o = O.OracleMap2D(typ)
            assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
            for k in range(seq.n):
                assert o.feed(seq.frame(k), seq.poses[k]) == accepted[k]
            assert state_record(o, typ, 6) == rec["tiles"], "oracle != cv2 for %s type %d" % (name, typ)
            assert sha(o.get_image()[0]) == rec["mosaic"]
            O.set_f32_mode(0)
            o = O.OracleMap2D(typ)
            assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
            for k in range(seq.n):
                o.feed(seq.frame(k), seq.poses[k])
            img2, org2 = o.get_image()
            out["oracle249"]["%s/type%d" % (name, typ)] = {"tiles": state_record(o, typ, 6), "mosaic": sha(img2),
                                                           "mosaic_origin": list(org2), "stats": o.stats()}
    # ---- Map2DRender (type 4): the batch blender.  "render/cv2": the oracle with cv2 4.x's float association, recorded only
    # after its weighted-sum blends (1, 2) have been found equal to the REAL cv2.detail_MultiBandBlender fed with the same warped
    # frames, right here; "render/oracle249": the oracle's default (2.4.9) association, what the CUDA path defaults to.
    out["render"] = {}
    for name, kw in SEQS.items():
        seq = synth.Sequence(**kw)
        frames = seq.frames()
        for blend in (0, 1, 2):
            for mode, fam in ((1, "cv2"), (0, "oracle249")):
                O.set_f32_mode(mode)
                o = O.OracleMap2D(O.TYPE_RENDER, render_blend=blend)
                assert o.prepare(seq.plane, seq.camera, seq.prepare_poses)
                rc, res = o.render_frames(frames, seq.poses)
                assert rc == 0
                r16, mask, nb, org = o.render_get()
                if mode == 1 and blend in (1, 2):
                    b = cv2.detail_MultiBandBlender(0, nb, cv2.CV_32F if blend == 1 else cv2.CV_16S)
                    b.prepare((0, 0, r16.shape[1], r16.shape[0]))
                    for i in range(seq.n):
                        w = o.render_warped(i)
                        if w is not None:
                            b.feed(w[0], w[1], w[2])
                    ref, ref_mask = b.blend(None, None)
                    assert np.array_equal(ref, r16) and np.array_equal(ref_mask, mask), "render oracle != cv2 blender (%s blend %d)" % (name, blend)
                out["render"]["%s/blend%d/%s" % (name, blend, fam)] = {
                    "accepted": [int(v) for v in res], "bands": nb, "origin": list(org), "shape": list(r16.shape),
                    "result": sha(r16), "mask": sha(mask), "image8": sha(np.clip(r16, 0, 255).astype(np.uint8))}
    O.set_f32_mode(0)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)
    print("wrote golden.json (%d bytes), kat.npz (%d bytes)" % (os.path.getsize(os.path.join(HERE, "golden.json")),
                                                               os.path.getsize(os.path.join(HERE, "kat.npz"))))


if __name__ == "__main__":
    main()
