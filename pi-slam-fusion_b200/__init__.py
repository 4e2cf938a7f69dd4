"""B200-native Map2DFusion feed() hot path: CUDA kernels + C-ABI (csrc/, libmap2d_b200.so) and the host-side
mirror of the reference's Map2D plugin interface (map2d.py).  See DESIGN.md / INTEGRATION.md."""
__all__ = ["map2d", "synth"]
