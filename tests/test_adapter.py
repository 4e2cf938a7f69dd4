"""include/Map2DB200.h (the C++ `class Map2DB200 : public Map2D` adapter of INTEGRATION.md) compiles against
stand-ins of the reference headers, links to libmap2d_b200.so, and behaves like the reference driver expects."""
import os
import re
import subprocess

import pytest

import pi_slam_fusion_b200.map2d as m2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "adapter_stub")


def build(tmp_path):
    exe = str(tmp_path / "adapter_stub")
    libdir = os.path.dirname(m2d.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-Wall", "-I", STUB, "-I", os.path.join(ROOT, "include"),
                           os.path.join(STUB, "main.cpp"), "-o", exe, "-L", libdir, "-lmap2d_b200", "-Wl,-rpath," + libdir])
    return exe


def run(exe, *args):
    out = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return dict(kv.split("=") for kv in out.stdout.strip().splitlines()[-1].split())


def test_adapter_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    exe = build(tmp_path)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("covered by the gpu test")
    r = run(exe)
    assert r["handle"] == "0" and r["prepared"] == "0" and r["fed"] == "0" and r["saved"] == "0"  # no silent CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("typ", [1, 2, 3])
def test_adapter_drives_the_gpu_path(tmp_path, typ):
    exe = build(tmp_path)
    png = str(tmp_path / "out.png")
    r = run(exe, str(typ), png)
    assert r["handle"] == "1" and r["prepared"] == "1" and r["fed"] == "4" and r["oblique_accepted"] == "0" and r["saved"] == "1"
    w, h = map(int, re.match(r"(\d+)x(\d+)", r["image"]).groups())
    assert w > 0 and w % 256 == 0 and h % 256 == 0 and os.path.getsize(png) > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("typ", [1, 3])
def test_adapter_thread_mode_uses_the_ingest_queue(tmp_path, typ):
    """thread=true: feed() only enqueues (so even the oblique frame is 'accepted', like the reference's enqueue),
    save()/getImage() drain the queue first.  Map2DPrepare::_frames is the worker's queue (Map2D.cpp:42,
    Map2DCPU.cpp:384-413), so in thread mode the prepare-frames 0..3 ARE rendered, before anything feed() adds; in
    thread=false they are not (only feed() renders)."""
    exe = build(tmp_path)
    a, b, c, d = (str(tmp_path / n) for n in ("sync_0to7.png", "thread_4to7.png", "thread_none.png", "sync_0to3.png"))
    r0 = run(exe, str(typ), a, "0", "8", "0")       # synchronous: feed frames 0..7 explicitly
    r1 = run(exe, str(typ), b, "1", "4", "4")       # threaded: prepare(0..3) seeds the queue, feed adds 4..7
    assert r1["handle"] == "1" and r1["prepared"] == "1" and r1["fed"] == "4" and r1["oblique_accepted"] == "1" and r1["saved"] == "1"
    assert r1["image"] == r0["image"] and r1["queue"] == "0"
    assert open(a, "rb").read() == open(b, "rb").read(), "thread mode must render the prepare-frames first, then the fed ones"
    r2 = run(exe, str(typ), c, "1", "0", "0")       # threaded, nothing fed: the mosaic is the prepare-frames alone
    r3 = run(exe, str(typ), d, "0", "4", "0")
    assert r2["saved"] == "1" and open(c, "rb").read() == open(d, "rb").read()
    r4 = run(exe, str(typ), str(tmp_path / "sync_none.png"), "0", "0", "0")   # thread=false never renders the prepare set
    assert r4["prepared"] == "1" and r4["saved"] == "0" and r4["image"] == "0x0"


@pytest.mark.gpu
def test_adapter_type_render(tmp_path):
    """Map2D::create(TypeRender, thread): thread=true renders the prepare-frames as ONE batch at prepare() (the reference's
    worker takes the whole queue through renderFrames and stops, Map2DRender.cpp:812-829, 758) and writes result.png next to
    the process (reference: result.jpg); feed() only queues.  thread=false: feed() is false, nothing is ever rendered."""
    exe = build(tmp_path)
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        r = run(exe, "4", str(tmp_path / "render.png"), "1", "4", "4")
        assert r["handle"] == "1" and r["prepared"] == "1" and r["fed"] == "4" and r["oblique_accepted"] == "1" and r["saved"] == "1"
        w, h = map(int, re.match(r"(\d+)x(\d+)", r["image"]).groups())
        assert w > 0 and w % 256 == 0 and h % 256 == 0
        assert open(str(tmp_path / "render.png"), "rb").read() == open(str(tmp_path / "result.png"), "rb").read()
        r = run(exe, "4", str(tmp_path / "none.png"), "0", "4", "0")
        assert r["prepared"] == "1" and r["fed"] == "0" and r["saved"] == "0" and r["image"] == "0x0"
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("threaded", ["0", "1"])
def test_adapter_drives_two_gpus_in_one_process(tmp_path, typ, threaded):
    """Multi-GPU behind the boundary: the unmodified host loop, ONE process, `Map2DB200(type, thread, {0, 1})` ->
    m2d_create_multi.  Tiles are sharded over the two devices, every device sees every frame, save() gathers over NVLink.
    The PNG must equal the single-GPU one byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = build(tmp_path)
    a, b = str(tmp_path / "one.png"), str(tmp_path / "two.png")
    r1 = run(exe, str(typ), a, threaded, "8", "0", "1")
    r2 = run(exe, str(typ), b, threaded, "8", "0", "2")
    assert r2["handle"] == "1" and r2["prepared"] == "1" and r2["fed"] == "8" and r2["saved"] == "1" and r2["image"] == r1["image"]
    assert open(a, "rb").read() == open(b, "rb").read()
