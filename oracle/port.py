"""ctypes wrapper of oracle/liboracle_port.so (rt_oracle.cpp) — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs, nowhere else.  Takes the same flattened rt_scene_desc pointer the CUDA
library takes (from raytracingoneweekendapplication_b200.capi.Scene).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def ref_driver_path() -> str:
    return os.path.join(_HERE, "_ref", "ref_driver")


def have_ref_driver() -> bool:
    return os.path.exists(ref_driver_path())


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "liboracle_port.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", _HERE, "port"])
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.oracle_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, vp, vp]
    lib.oracle_primary.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.oracle_hit.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    lib.oracle_scatter.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.oracle_texture.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    _lib = lib
    return lib


def _ptr(scene):
    p = getattr(scene, "desc_ptr", scene)
    return C.cast(p, C.c_void_p)


def render(scene, width, height, spp, depth=50, seed=1, spp_begin=0, threads=0, want_var=False):
    lib = load()
    rgb = np.zeros((height, width, 3), dtype=np.float64)
    var = np.zeros((height, width, 3), dtype=np.float64) if want_var else None
    rc = lib.oracle_render(_ptr(scene), width, height, spp, spp_begin, depth, seed, threads, rgb.ctypes.data,
                           var.ctypes.data if want_var else None)
    if rc != 0:
        raise RuntimeError("oracle_render failed")
    return (rgb, var) if want_var else rgb


def primary(scene, width, height) -> dict:
    lib = load()
    n = width * height
    out = {"prim_id": np.zeros(n, np.int32), "t": np.zeros(n), "normal": np.zeros((n, 3)), "point": np.zeros((n, 3)), "uv": np.zeros((n, 2))}
    rc = lib.oracle_primary(_ptr(scene), width, height, *[out[k].ctypes.data for k in ("prim_id", "t", "normal", "point", "uv")])
    if rc != 0:
        raise RuntimeError("oracle_primary failed")
    return out


def hit(scene, rays) -> dict:
    lib = load()
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 9)
    n = rays.shape[0]
    out = {"prim_id": np.zeros(n, np.int32), "t": np.zeros(n), "normal": np.zeros((n, 3)), "uv": np.zeros((n, 2))}
    rc = lib.oracle_hit(_ptr(scene), n, rays.ctypes.data, *[out[k].ctypes.data for k in ("prim_id", "t", "normal", "uv")])
    if rc != 0:
        raise RuntimeError("oracle_hit failed")
    return out


def scatter(scene, material, records, uniforms) -> np.ndarray:
    lib = load()
    records = np.ascontiguousarray(records, dtype=np.float64).reshape(-1, 16)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1, 4)
    out = np.zeros((records.shape[0], 16))
    rc = lib.oracle_scatter(_ptr(scene), material, records.shape[0], records.ctypes.data, uniforms.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_scatter failed")
    return out


def texture(scene, tex, uvp) -> np.ndarray:
    lib = load()
    uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
    out = np.zeros((uvp.shape[0], 3))
    rc = lib.oracle_texture(_ptr(scene), tex, uvp.shape[0], uvp.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_texture failed")
    return out
