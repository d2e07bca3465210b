"""vkhr_b200 -- B200-native strand voxelisation, a drop-in for one path of CaffeineViking/vkhr.

The path: ``HairStyle::voxelize_vertices`` / ``voxelize_segments`` -> ``HairStyle::Volume``
(reference src/vkhr/scene_graph/hair_style.cc:257-342), rebuilt as hand-written sm_100a
CUDA behind the C ABI of ``include/vkhr_b200.h``.  Importing this package loads
``vkhr_b200/lib/libvkhr_b200.so``; if the library is absent and cannot be built the
import fails -- there is no CPU fallback.
"""
from . import capi                                   # noqa: F401  (loads the native library, or raises)
from .capi import (BRICK8_SPLIT, DOWNSAMPLE_MAX, DOWNSAMPLE_MEAN, DOWNSAMPLE_MIN, DOWNSAMPLE_SUM, INDEX_EXACT, NORMALIZE,
                   STRATEGY_BRICK8, STRATEGY_COUNT32, STRATEGY_PACKED8, VkhrB200Error)
from .hair_style import AABB, HairStyle, Volume
from .voxelizer import Voxelizer, default_voxelizer

__all__ = ["AABB", "HairStyle", "Volume", "Voxelizer", "default_voxelizer", "VkhrB200Error",
           "INDEX_EXACT", "NORMALIZE", "BRICK8_SPLIT", "STRATEGY_BRICK8", "STRATEGY_COUNT32", "STRATEGY_PACKED8",
           "DOWNSAMPLE_MAX", "DOWNSAMPLE_MEAN", "DOWNSAMPLE_SUM", "DOWNSAMPLE_MIN"]
__version__ = "0.1.0"
