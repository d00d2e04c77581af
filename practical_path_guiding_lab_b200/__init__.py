"""B200-native SD-tree hot path of the Mitsuba 3 "Practical Path Guiding" lab.

The product is libsdtree.so (hand-written sm_100a CUDA behind the C ABI of
include/sdtree.h); this package is the thin Python side: the ctypes binding (_lib),
the handle wrapper (sdtree.SDTree) and the drop-in integrator (integrator).
"""
from .sdtree import SDTree, SDTreeError, NPZ_KEYS  # noqa: F401
