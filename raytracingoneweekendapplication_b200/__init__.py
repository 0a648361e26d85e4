"""raytracingoneweekendapplication_b200 — B200-native path-tracing core behind the scene API of
grahamstockton87/RayTracingOneWeekendApplication.

  csrc/   hand-written sm_100a CUDA kernels + the C ABI of include/rt_b200.h (librt_b200.so)
  host/   C++ host mirror of the reference's scene-construction API (flatten -> C ABI)
  capi.py ctypes binding used by the tests and bench.py
"""
from . import capi  # noqa: F401
from .capi import Context, RtError, Scene, scene_names  # noqa: F401
