"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs, never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_REPO, "rust-raytracer_b200"))
import rtb200 as R  # struct definitions (include/rtb200.h mirrors) only  # noqa: E402

# RTB200_ORACLE_LIB: load another build of the same sources instead (tests/test_oracle_ubsan.py loads a UBSan build)
LIB_PATH = os.environ.get("RTB200_ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")


class oracle_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples", C.c_uint64), ("draws", C.c_uint64), ("hits", C.c_uint64 * 5),
                ("term_sky", C.c_uint64), ("term_absorbed", C.c_uint64), ("term_depth", C.c_uint64), ("term_light", C.c_uint64),
                ("texture_oob", C.c_uint64), ("path_len_hist", C.c_uint64 * 64), ("render_ms", C.c_double),
                ("threads", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("hits", "path_len_hist")}
        d["hits"] = list(self.hits)
        d["path_len_hist"] = list(self.path_len_hist)
        return d


_lib = None


def build(force: bool = False):
    """(Re)build liboracle.so when it is missing or older than its sources / include/rtb200.h (the structs it reads)."""
    if os.environ.get("RTB200_ORACLE_LIB"):
        return
    deps = [os.path.join(_HERE, f) for f in ("rt_oracle_capi.cpp", "rt_oracle.hpp", "Makefile")] + [os.path.join(_REPO, "include", "rtb200.h")]
    stale = not os.path.exists(LIB_PATH) or any(os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_render.argtypes = [C.POINTER(R.rt_scene), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(oracle_stats)]
        L.oracle_sample.argtypes = [C.POINTER(R.rt_scene), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_camera_new.argtypes = [C.POINTER(R.rt_camera_params), C.POINTER(R.rt_camera)]
        L.oracle_get_ray.argtypes = [C.POINTER(R.rt_camera), C.c_double, C.c_double, C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3)]
        L.oracle_ray_at.argtypes = [C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3), C.c_double, C.POINTER(R.rt_vec3)]
        L.oracle_sphere_hit.argtypes = [C.POINTER(R.rt_vec3), C.c_double, C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3), C.c_double, C.c_double,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_refract.argtypes = [C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3), C.c_double, C.POINTER(R.rt_vec3)]
        L.oracle_reflect.argtypes = [C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3)]
        L.oracle_reflectance.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
        L.oracle_ray_color.argtypes = [C.POINTER(R.rt_scene), C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3), C.c_uint64, C.c_uint64, C.POINTER(C.c_float)]
        L.oracle_find_lights.argtypes = [C.POINTER(R.rt_scene), C.POINTER(C.c_int32), C.c_uint32]
        L.oracle_rng.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.oracle_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_quantise.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_texture_albedo.argtypes = [C.POINTER(R.rt_image), C.c_double, C.c_double, C.c_double, C.POINTER(C.c_float)]
        L.oracle_p3_ops.argtypes = [C.POINTER(R.rt_vec3), C.POINTER(R.rt_vec3), C.c_double, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def render(scene: "R.Scene", linear: bool = True, rgb8: bool = True, y0: int = 0, y1: int = 0, threads: int = 0):
    """Render rows [y0,y1) with the oracle. Returns (linear float32 [h,w,3] | None, rgb8 uint8 [h,w,3] | None, stats dict)."""
    h, w = scene.c.height, scene.c.width
    lin = np.zeros((h, w, 3), dtype=np.float32) if linear else None
    img = np.zeros((h, w, 3), dtype=np.uint8) if rgb8 else None
    st = oracle_stats()
    rc = lib().oracle_render(C.byref(scene.c), lin.ctypes.data if linear else None, img.ctypes.data if rgb8 else None, y0, y1, threads, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    return lin, img, st.as_dict()


def sample(scene: "R.Scene", x: int, y: int, s: int):
    out = (C.c_float * 3)()
    rays = C.c_uint64(); draws = C.c_uint64()
    lib().oracle_sample(C.byref(scene.c), x, y, s, out, C.byref(rays), C.byref(draws))
    return np.array(out[:], dtype=np.float32), rays.value, draws.value
