"""B200-native time-of-flight ray-tracing engine (drop-in for the hot path of
edudscrc/ray-tracing-ultrasound: main_rt.py shoot_rays + table assembly).

The compute path is hand-written CUDA for sm_100a behind the C ABI of
include/tofrt.h (libtofrt.so, built in-tree by `_build.build()`); this package
is the host-side mirror of the reference interface.  There is no CPU fallback.
"""
from . import reference_setup
from ._capi import TofrtError
from .engine import TofEngine, TofTable, pinned_empty

__all__ = ["TofEngine", "TofTable", "TofrtError", "pinned_empty", "reference_setup"]
