#!/usr/bin/env python3
"""C1 (test_scene: textures, light, hollow glass, sky texture) GPU vs oracle: where do pixels differ, and does the
device u_v_from_sphere_hit_point equal the oracle's bit for bit? (run under gpurun)"""
import ctypes as C, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'rust-raytracer_b200')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import numpy as np
import rtb200 as R
from rtb200 import scenes
import oracle_py as O

rng = np.random.default_rng(7)
n = 1 << 20
hp = rng.normal(size=(n, 3)) * rng.uniform(0.1, 50.0, size=(n, 1))
hp = np.ascontiguousarray(hp)
a = np.zeros((n, 2)); b = np.zeros((n, 2))
rc = R.lib().rtb200_probe_sphere_uv(hp.ctypes.data_as(C.POINTER(C.c_double)), n, a.ctypes.data_as(C.POINTER(C.c_double)))
assert rc == 0, R.lib().rtb200_last_error()
O.lib().oracle_sphere_uv(hp.ctypes.data_as(C.POINTER(C.c_double)), n, b.ctypes.data_as(C.POINTER(C.c_double)))
print("sphere_uv: device != oracle(restated atan2) in", int((a != b).any(axis=1).sum()), "of", n, flush=True)
O.lib().oracle_set_atan2_mode(0)
O.lib().oracle_sphere_uv(hp.ctypes.data_as(C.POINTER(C.c_double)), n, b.ctypes.data_as(C.POINTER(C.c_double)))
O.lib().oracle_set_atan2_mode(1)
print("sphere_uv: device != oracle(libm atan2) in", int((a != b).any(axis=1).sum()), "of", n, flush=True)

sc = scenes.scene("C1")
for mode in (1, 0):
    O.lib().oracle_set_atan2_mode(mode)
    lin_o, img_o, st_o = O.render(sc)
    lin_g, st_g = R.render_linear(sc)
    d = np.abs(lin_g - lin_o).max(axis=2)
    ys, xs = np.nonzero(d > 0)
    print(f"C1 atan2_mode={mode}: differing pixels {len(ys)} of {d.size}; rays gpu {st_g['rays']} oracle {st_o['rays']}; max|d| {d.max():.3e}", flush=True)
    for y, x in list(zip(ys, xs))[:12]:
        print("   pixel", x, y, lin_g[y, x], lin_o[y, x], flush=True)
O.lib().oracle_set_atan2_mode(1)
