"""raytracingincuda_b200 -- B200-native `render` hot path of jilinzheng/RaytracingInCUDA.

The product is the C-ABI library ``librt_b200.so`` (include/rt_b200.h) and the drop-in CLI
``bin/b200-raytrace``; this package is the Python host glue over that ABI used by the tests,
``bench.py`` and the multi-GPU (one process per GPU) driver.  There is no CPU fallback: importing
:mod:`raytracingincuda_b200.api` raises if the library has not been built.
"""
from .api import (Camera, Camera64, Opts, Renderer, RtError, Slot, Slot64, SLOT_DTYPE, SLOT64_DTYPE, Stats,
                  camera, lib, load_scene, num_chunks, partition_samples, partition_rows, ppm_quantise, ppm_write, save_scene, scene,
                  scene_scaled)

__all__ = ["Camera", "Camera64", "Opts", "Renderer", "RtError", "Slot", "Slot64", "SLOT_DTYPE", "SLOT64_DTYPE",
           "Stats", "camera", "lib", "load_scene", "num_chunks", "partition_samples", "partition_rows", "ppm_quantise",
           "ppm_write", "save_scene", "scene", "scene_scaled"]
