"""The C-ABI library loads on a box without a GPU, exports every symbol include/map2d_b200.h declares, agrees with
the Python mirror on struct layout, and FAILS LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

import pi_slam_fusion_b200.map2d as m2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "map2d_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m2d_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(m2d.LIB_PATH), "run `python __graft_entry__.py` (build()) first"
    m2d.lib()


def test_every_declared_symbol_is_exported():
    L = m2d.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "symbol %s declared in map2d_b200.h is not exported" % s
    assert sorted(m2d.EXPORTS) == syms, "python mirror out of sync with the header"


def test_only_m2d_symbols_have_default_visibility_prefix():
    out = subprocess.run(["nm", "-D", "--defined-only", m2d.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(s in exported for s in declared_symbols())


def test_config_struct_layout_matches():
    c = m2d.Config()
    for f, _ in m2d.Config._fields_:
        setattr(c, f, 7)
    m2d.lib().m2d_config_default(C.byref(c))
    assert c.scale == 1.0 and c.band_number == 5 and c.shard_count == 1 and c.resolution == 0.0
    assert c.weight_type == 0 and c.force_float == 0 and c.collect_stats == 0 and c.batch_frames == 0
    # header struct: 2 doubles + 12 ints
    assert C.sizeof(m2d.Config) == 2 * 8 + 12 * 4
    assert C.sizeof(m2d.Stats) == 8 * (3 + 3 * m2d.MAX_LEVELS + 1 + m2d.MAX_LEVELS)  # ... + need_px[MAX_LEVELS]


def test_unsupported_types_and_arguments():
    L = m2d.lib()
    h = C.c_void_p()
    cfg = m2d.default_config()
    assert L.m2d_create(0, C.byref(cfg), C.byref(h)) == -5 and not h.value      # NoType
    assert L.m2d_create(4, C.byref(cfg), C.byref(h)) == -5 and not h.value      # TypeRender: out of scope
    ff = m2d.default_config(force_float=1)
    assert L.m2d_create(3, C.byref(ff), C.byref(h)) == -5 and not h.value       # ForceFloat unsupported
    assert m2d.Map2D.create(0) is None                                           # null SPtr in the reference
    assert L.m2d_sync(None) == -1 and L.m2d_queue_size(None) == 0


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    cfg = m2d.default_config()
    assert m2d.lib().m2d_create(1, C.byref(cfg), C.byref(h)) == -3 and not h.value  # M2D_ERR_CUDA, loudly
    with pytest.raises(RuntimeError):
        m2d.Map2D.create(3, thread=False)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pi-slam-fusion_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "map2d_oracle" not in txt, fn
