"""Pins the CPU oracle against every known-answer test the reference's own suite holds for the hot path
(SURVEY.md §8c). CPU only."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_py as O
import rtb200 as R

V = R.vec3


def test_sphere_hit_kat(kat):
    k = kat["sphere_hit"]   # sphere.rs:81-88
    hit = C.c_int32(); t = C.c_double(); p = R.rt_vec3(); n = R.rt_vec3(); ff = C.c_int32(); u = C.c_double(); v = C.c_double()
    O.lib().oracle_sphere_hit(V(k["center"]), k["radius"], V(k["origin"]), V(k["dir"]), k["t_min"], math.inf,
                              C.byref(hit), C.byref(t), C.byref(p), C.byref(n), C.byref(ff), C.byref(u), C.byref(v))
    assert hit.value == 1 and t.value == k["t"]
    assert p.tup() == (0.0, 0.0, -1.0) and n.tup() == (0.0, 0.0, -1.0) and ff.value == 1


def test_sphere_hit_edges():
    L = O.lib()
    hit = C.c_int32(); t = C.c_double(); p = R.rt_vec3(); n = R.rt_vec3(); ff = C.c_int32(); u = C.c_double(); v = C.c_double()
    call = lambda c, r, o, d, tmin, tmax: L.oracle_sphere_hit(V(c), r, V(o), V(d), tmin, tmax, C.byref(hit), C.byref(t), C.byref(p), C.byref(n), C.byref(ff), C.byref(u), C.byref(v))
    # origin inside: near root negative, far root taken, normal flipped to face the ray (sphere.rs:55-68)
    call([0, 0, 0], 1.0, [0, 0, 0], [0, 0, 1], 0.001, math.inf)
    assert hit.value == 1 and t.value == 1.0 and ff.value == 0 and n.tup() == (-0.0, -0.0, -1.0)
    # t_max shrinks below both roots -> miss (hit_world's closest_so_far, raytracer.rs:53)
    call([0, 0, 0], 1.0, [0, 0, -5], [0, 0, 1], 0.001, 3.0)
    assert hit.value == 0
    # negative radius flips the normal (hollow shell, test_scene.json:137)
    call([0, 0, 0], -1.0, [0, 0, -5], [0, 0, 1], 0.001, math.inf)
    assert hit.value == 1 and t.value == 4.0 and ff.value == 0
    # tangent ray: discriminant == 0 still counts (>= 0.0, sphere.rs:53)
    call([0, 1, 0], 1.0, [0, 0, -5], [0, 0, 1], 0.001, math.inf)
    assert hit.value == 1 and t.value == 5.0
    # un-normalised direction: t scales by 1/|d|
    call([0, 0, 0], 1.0, [0, 0, -5], [0, 0, 2], 0.0, math.inf)
    assert hit.value == 1 and t.value == 2.0


def test_refract_reflectance_kat(kat):
    k = kat["refract"]      # materials.rs:157-165
    out = R.rt_vec3()
    O.lib().oracle_refract(V(k["uv"]), V(k["n"]), k["eta"], C.byref(out))
    assert out.tup() == tuple(float(x) for x in k["expect"])
    k = kat["reflectance"]  # materials.rs:167-174
    r = C.c_double()
    O.lib().oracle_reflectance(k["cosine"], k["ref_idx"], C.byref(r))
    assert r.value == k["expect"]


def _empty_scene(sky_mode):
    sc = R.Scene()
    sc.c.width, sc.c.height, sc.c.samples_per_pixel, sc.c.max_depth = 80, 60, 1, 2
    sc.c.camera = R.camera_from_params([0, 0, -3], [0, 0, 0], [0, 1, 0], 20.0, 1.333)
    sc.c.sky.mode = sky_mode
    sc.c.n_spheres = 0
    return sc


def test_ray_color_sky_kat(kat):
    k = kat["ray_color_sky"]   # raytracer.rs:167-189: exact f32 equality
    out = (C.c_float * 3)()
    O.lib().oracle_ray_color(C.byref(_empty_scene(R.RT_SKY_GRADIENT).c), V(k["origin"]), V(k["dir"]), 2, 2, out)
    assert [out[0], out[1], out[2]] == [np.float32(x) for x in k["expect"]]
    O.lib().oracle_ray_color(C.byref(_empty_scene(R.RT_SKY_NONE).c), V(k["origin"]), V(k["dir"]), 2, 2, out)
    assert list(out) == [0.0, 0.0, 0.0]     # sky: None -> black (raytracer.rs:138-140)
    O.lib().oracle_ray_color(C.byref(_empty_scene(R.RT_SKY_GRADIENT).c), V(k["origin"]), V(k["dir"]), 2, 0, out)
    assert list(out) == [0.0, 0.0, 0.0]     # depth 0 -> black (raytracer.rs:80-82)


def test_camera_kat(kat):
    k = kat["camera_llc"]   # camera.rs:87-103
    p = R.rt_camera_params(V(k["look_from"]), V(k["look_at"]), V(k["vup"]), k["vfov"], k["aspect"])
    cam = R.rt_camera()
    O.lib().oracle_camera_new(C.byref(p), C.byref(cam))
    assert cam.origin.tup() == (0.0, 0.0, 0.0)
    np.testing.assert_allclose(cam.lower_left_corner.tup(), k["lower_left_corner"], atol=k["tol"], rtol=0)
    k = kat["camera_get_ray"]   # camera.rs:105-122
    p = R.rt_camera_params(V(k["look_from"]), V(k["look_at"]), V(k["vup"]), k["vfov"], k["aspect"])
    O.lib().oracle_camera_new(C.byref(p), C.byref(cam))
    o = R.rt_vec3(); d = R.rt_vec3()
    O.lib().oracle_get_ray(C.byref(cam), k["u"], k["v"], C.byref(o), C.byref(d))
    assert o.tup() == tuple(float(x) for x in k["origin"])
    np.testing.assert_allclose(d.tup(), k["dir"], atol=k["tol"], rtol=0)


def test_ray_at_and_point3d_kat(kat):
    k = kat["ray_at"]   # ray.rs:52-63
    out = R.rt_vec3()
    O.lib().oracle_ray_at(V(k["origin"]), V(k["dir"]), k["t"], C.byref(out))
    np.testing.assert_allclose(out.tup(), k["expect"], atol=k["tol"], rtol=0)
    k = kat["point3d"]  # point3d.rs:196-272
    o = (C.c_double * 24)()
    O.lib().oracle_p3_ops(V(k["p"]), V(k["q"]), 2.0, o)
    np.testing.assert_allclose(o[0:3], k["add"], atol=k["tol"], rtol=0)
    np.testing.assert_allclose(o[3:6], k["sub"], atol=k["tol"], rtol=0)
    np.testing.assert_allclose(o[6:9], k["neg"], atol=k["tol"], rtol=0)
    assert abs(o[15] - k["dot"]) < k["tol"] and abs(o[16] - k["length_squared"]) < k["tol"]
    assert o[21] == 0.0
    O.lib().oracle_p3_ops(V([0, 0, 0]), V(k["q"]), 2.0, o)
    assert o[21] == 1.0   # near_zero (point3d.rs:266-272)


def test_find_lights_kat(kat):
    sc = R.Scene.from_config({
        "width": 8, "height": 8, "samples_per_pixel": 1, "max_depth": 2, "sky": {"texture": ""},
        "camera": {"look_from": {"x": 0, "y": 0, "z": 0}, "look_at": {"x": 0, "y": 0, "z": -1}, "vup": {"x": 0, "y": 1, "z": 0}, "vfov": 90.0, "aspect": 1.0},
        "objects": [{"center": {"x": 0, "y": 0, "z": -1}, "radius": 0.5, "material": {"Light": {}}},
                    {"center": {"x": 0, "y": 0, "z": -1}, "radius": 0.5, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}}]})
    idx = (C.c_int32 * 4)()
    assert O.lib().oracle_find_lights(C.byref(sc.c), idx, 4) == kat["find_lights"]["expect"] and idx[0] == 0   # raytracer.rs:231-248


def test_philox_kat(kat):
    for v in kat["philox4x32_10"]["vectors"]:
        ctr = (C.c_uint32 * 4)(*[int(x, 16) for x in v["ctr"]]); key = (C.c_uint32 * 2)(*[int(x, 16) for x in v["key"]])
        out = (C.c_uint32 * 4)()
        O.lib().oracle_philox(ctr, key, out)
        assert [f"{x:08x}" for x in out] == v["out"]


def test_rng_stream_contract():
    n = 4096
    a = (C.c_double * n)(); b = (C.c_double * n)()
    O.lib().oracle_rng(0x5EED, 7, 3, 0, n, a)
    O.lib().oracle_rng(0x5EED, 7, 3, 1, n, b)
    a = np.array(a[:]); b = np.array(b[:])
    assert a.min() >= 0.0 and a.max() < 1.0 and b.min() >= -1.0 and b.max() < 1.0        # point3d.rs:258-264 range check
    assert np.all(a * 2.0**53 == np.floor(a * 2.0**53))                                  # 53-bit grid (rand Standard)
    assert abs(a.mean() - 0.5) < 0.03 and abs(b.mean()) < 0.06
    # same u64 stream behind both conversions: b = 2*floor52(a) - 1
    assert np.all(np.abs((b + 1.0) / 2.0 - a) < 2.0**-52)
    c = (C.c_double * 8)()
    O.lib().oracle_rng(0x5EED, 7, 4, 0, 8, c)
    assert list(c) != list(a[:8])   # different sample -> different stream


def test_quantise_matches_palette_rounding():
    # palette 0.6 into_format::<u8>: round-half-even of min(x*255, 255); NaN/negative -> 0
    xs = np.array([0.0, 1.0, 4.0, 0.25, (0.5 / 255.0) ** 2, (1.5 / 255.0) ** 2, (2.5 / 255.0) ** 2, 1e-12], dtype=np.float32)
    out = np.zeros(len(xs), dtype=np.uint8)
    O.lib().oracle_quantise(xs.ctypes.data, len(xs), out.ctypes.data)
    exp = [int(np.rint(min(np.float32(np.sqrt(np.float32(x))) * np.float32(255.0), np.float32(255.0)))) for x in xs]
    assert list(out) == exp
    assert out[0] == 0 and out[1] == 255 and out[2] == 255 and out[3] == 128
