"""B200-native Map2DFusion feed() hot path: CUDA kernels + C-ABI (csrc/, libmap2d_b200.so) and the host-side
mirror of the reference's Map2D plugin interface.  See DESIGN.md / INTEGRATION.md.

    map2d     ctypes mirror of include/map2d_b200.h with the reference's method names (create / prepare / feed / save /
              queueSize), plus in-memory getters, the ingest queue, display tiles, checkpoints, the sharding hooks and
              the Map2DRender batch blender (render_frames / render_get)
    sharded   one process per GPU: tile ownership, frame delivery plans, in-place halo sampling, sharded save (torch.distributed)
    replay    Map2DFusion dataset format + headless replay driver
    synth     deterministic synthetic nadir surveys (tests, bench)
"""
__all__ = ["map2d", "synth", "sharded", "replay"]
