"""Pins oracle behaviours that the reference's own tests do not cover but its source fixes line by line
(texture addressing, sky texture addressing, light-test short-circuits, usize wrap, Metal absorption)."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
import rtb200 as R
from synth import base_config, _v


def _img(w, h):
    a = np.zeros((h, w, 3), np.uint8)
    a[..., 0] = (np.arange(w)[None, :] % 256); a[..., 1] = (np.arange(h)[:, None] % 256); a[..., 2] = 7
    return np.ascontiguousarray(a)


def _albedo(img, w, h, h_offset, u, v):
    im = R.rt_image(img.ctypes.data, w, h, img.size)
    out = (C.c_float * 3)()
    oob = O.lib().oracle_texture_albedo(C.byref(im), h_offset, u, v, out)
    return [out[0], out[1], out[2]], oob


def test_texture_get_albedo_addressing():
    """materials.rs:236-253: rot = u + h_offset (wrapped once when > 1), x = floor(rot*W), y = floor((1-v)*(H-1)), rgb/255."""
    w, h = 64, 32
    img = _img(w, h)
    f = lambda x: np.float32(x) / np.float32(255.0)
    assert _albedo(img, w, h, 0.0, 0.5, 0.5) == ([f(32), f(15), f(7)], 0)            # x = 32, y = floor(15.5) = 15
    assert _albedo(img, w, h, 0.75, 0.5, 1.0) == ([f(16), f(0), f(7)], 0)            # rot = 1.25 -> 0.25 -> x = 16; v = 1 -> y = 0
    assert _albedo(img, w, h, 0.0, 0.0, 0.0) == ([f(0), f(31), f(7)], 0)             # v = 0 -> y = H-1
    rgb, oob = _albedo(img, w, h, 0.0, 1.0, 0.0)                                     # rot == 1.0 is NOT wrapped: x = W on the last row
    assert oob == 1                                                                  # the reference indexes out of bounds (panic); oracle clamps and counts


def test_sky_texture_addressing_and_black_sky():
    """raytracer.rs:149-159: x = (u*(W-1)) as usize, y = ((1-t)*(H-1)) as usize, colour = 0.7 * texel / 255."""
    w, h = 16, 8
    img = _img(w, h)
    sc = R.Scene.from_config(base_config(8, 8, 1, 2, [], sky="gradient"))
    sc.c.sky.mode = R.RT_SKY_TEXTURE
    sc.c.sky.tex = R.rt_image(img.ctypes.data, w, h, img.size)
    out = (C.c_float * 3)()
    O.lib().oracle_ray_color(C.byref(sc.c), R.vec3([0, 0, 0]), R.vec3([0, 1, 0]), 2, 2, out)      # straight up: t = 1, u = 0.5
    x, y = int(np.float32(0.5) * np.float32(w - 1)), 0
    exp = [np.float32(0.7) * np.float32(img[y, x, c]) / np.float32(255.0) for c in range(3)]
    assert [out[0], out[1], out[2]] == exp
    O.lib().oracle_ray_color(C.byref(sc.c), R.vec3([0, 0, 0]), R.vec3([1, -1, 0]), 2, 2, out)     # down-right
    ud = np.array([1, -1, 0]) / np.sqrt(2.0)
    t = np.float32(0.5) * (np.float32(ud[1]) + np.float32(1.0)); u = np.float32(0.5) * (np.float32(ud[0]) + np.float32(1.0))
    x, y = int(u * np.float32(w - 1)), int((np.float32(1.0) - t) * np.float32(h - 1))
    exp = [np.float32(0.7) * np.float32(img[y, x, c]) / np.float32(255.0) for c in range(3)]
    assert [out[0], out[1], out[2]] == exp


def _one_sphere(material, depth, lights=()):
    objs = [{"center": _v(0, 0, -3), "radius": 1.0, "material": material}] + [{"center": _v(*p), "radius": 0.5, "material": {"Light": {}}} for p in lights]
    return R.Scene.from_config(base_config(8, 8, 1, depth, objs, sky="none", look_from=(0, 0, 0), look_at=(0, 0, -1), vfov=40.0))


def test_light_hit_returns_white_and_absorbing_metal_returns_black():
    out = (C.c_float * 3)()
    sc = _one_sphere({"Light": {}}, 5)
    O.lib().oracle_ray_color(C.byref(sc.c), R.vec3([0, 0, 0]), R.vec3([0, 0, -1]), 5, 5, out)
    assert list(out) == [1.0, 1.0, 1.0]                       # `None => albedo` (raytracer.rs:124), white (materials.rs:67)
    # a Lambertian in a lightless, skyless scene is black at any depth
    sc = _one_sphere({"Lambertian": {"albedo": [0.9, 0.9, 0.9]}}, 5)
    O.lib().oracle_ray_color(C.byref(sc.c), R.vec3([0, 0, 0]), R.vec3([0, 0, -1]), 5, 5, out)
    assert list(out) == [0.0, 0.0, 0.0]


def test_light_test_depth_rule_and_usize_wrap():
    """`depth > max_depth - 2` (raytracer.rs:101): shadow rays only from the first two path levels; for max_depth < 2 the
    usize subtraction wraps (release build) and the test is false, so no shadow ray is ever cast."""
    lin = {}
    for depth in (1, 2, 3):
        sc = _one_sphere({"Lambertian": {"albedo": [0.8, 0.8, 0.8]}}, depth, lights=[(0, 3, -3)])
        sc.c.samples_per_pixel = 64
        l, _, st = O.render(sc)
        lin[depth] = (l, st)
    assert lin[1][1]["rays"] == lin[1][1]["samples"]           # max_depth 1: wrap -> never a shadow ray
    assert lin[2][1]["rays"] > lin[2][1]["samples"]            # max_depth 2: both levels may cast shadow rays
    assert lin[1][0].max() == 0.0 or lin[1][0].max() == 1.0    # only direct light hits (white) or black
    assert lin[2][0].max() > 0.0


def test_draw_order_is_scatter_then_light_test():
    """With one light every non-absorbed vertex consumes exactly one extra uniform after its scatter draws
    (raytracer.rs:86 precedes :100): the draw count of a light scene exceeds the lightless one accordingly."""
    objs = [{"center": _v(0, -100.5, -3), "radius": 100.0, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}}]
    a = R.Scene.from_config(base_config(16, 12, 4, 3, objs, sky="gradient", look_from=(0, 1, 2), look_at=(0, 0, -3), vfov=50.0))
    b = R.Scene.from_config(base_config(16, 12, 4, 3, objs + [{"center": _v(50, 80, -3), "radius": 0.1, "material": {"Light": {}}}], sky="gradient",
                                        look_from=(0, 1, 2), look_at=(0, 0, -3), vfov=50.0))
    _, _, sa = O.render(a); _, _, sb = O.render(b)
    assert sb["draws"] > sa["draws"] and sb["rays"] >= sa["rays"]


def test_restated_atan2_tracks_the_host_libm_and_does_not_move_a_texel():
    """f64::atan2 (sphere.rs:38) is libm-defined in the reference; oracle and kernel share one explicit algorithm instead
    (rto::rt_atan2). It stays within 2 ulp of this host's libm, agrees on every special value, and the C1 frame (the only
    BASELINE config with textures) is identical whichever of the two the oracle uses."""
    import ctypes as C
    from rtb200 import scenes
    L = O.lib()
    rng = np.random.default_rng(1)
    n = 200_000
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    y = np.ascontiguousarray(v[:, 0]); x = np.ascontiguousarray(v[:, 2])
    a = np.zeros(n); b = np.zeros(n)
    P = C.POINTER(C.c_double)
    L.oracle_atan2(y.ctypes.data_as(P), x.ctypes.data_as(P), n, 1, a.ctypes.data_as(P))
    L.oracle_atan2(y.ctypes.data_as(P), x.ctypes.data_as(P), n, 0, b.ctypes.data_as(P))
    assert np.abs(a.view(np.int64) - b.view(np.int64)).max() <= 2
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-310, -1e-310, 1e300, -1e300, 0.5, 2.0, -3.0, 1e-30])
    yy, xx = [np.ascontiguousarray(t.ravel()) for t in np.meshgrid(sp, sp)]
    a = np.zeros(len(yy)); b = np.zeros(len(yy))
    L.oracle_atan2(yy.ctypes.data_as(P), xx.ctypes.data_as(P), len(yy), 1, a.ctypes.data_as(P))
    L.oracle_atan2(yy.ctypes.data_as(P), xx.ctypes.data_as(P), len(yy), 0, b.ctypes.data_as(P))
    same = (a == b) | (np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= 1e-15 * np.abs(b))
    assert same.all() and np.array_equal(np.signbit(a[~np.isnan(a)]), np.signbit(b[~np.isnan(b)]))
    sc = scenes.scene("C1")
    try:
        L.oracle_set_atan2_mode(0)
        lin0, img0, st0 = O.render(sc)
    finally:
        L.oracle_set_atan2_mode(1)
    lin1, img1, st1 = O.render(sc)
    assert np.array_equal(lin0, lin1) and np.array_equal(img0, img1) and st0["rays"] == st1["rays"]
