"""cloudbrush_b200 -- B200-native overlap discovery + transitive reduction for CloudBrush.

Host-side mirror of the Brush stage API over the C ABI in include/cloudbrush_gpu.h.
"""
from .api import BrushGpu, BrushGpuError, BrushConfig  # noqa: F401
from . import stages  # noqa: F401

__all__ = ["BrushGpu", "BrushGpuError", "BrushConfig", "stages"]
