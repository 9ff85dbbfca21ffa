"""CPU-side checks of the C-ABI boundary: the library loads without a GPU, exports every symbol the header
declares, its host-side Camera::new agrees with the oracle bit for bit, shard arithmetic is consistent, and
the render entry points FAIL LOUDLY when no B200 is present (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_py as O
import rtb200 as R
from rtb200 import scenes


def _header_symbols(repo):
    txt = open(os.path.join(repo, "include", "rtb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(repo):
    L = R.lib()
    syms = _header_symbols(repo)
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(L, s), f"librtb200.so does not export {s}"
    assert sorted(R.ABI_SYMBOLS) == syms
    assert L.rtb200_abi_version() == 2


def test_struct_layouts_match_header():
    assert C.sizeof(R.rt_vec3) == 24 and C.sizeof(R.rt_camera) == 96 and C.sizeof(R.rt_sphere) == 64
    assert C.sizeof(R.rt_image) == 32 and C.sizeof(R.rt_sky) == 40
    assert C.sizeof(R.rt_scene) == 16 + 96 + 40 + 16 + 16 + 8
    assert C.sizeof(R.rt_options) == 32 and C.sizeof(R.rt_stats) == 104 and C.sizeof(R.rt_kernel_info) == 140


@pytest.mark.parametrize("args", [
    ([0, 0, 0], [0, 0, -1], [0, 1, 0], 90.0, 800.0 / 600.0),      # camera.rs:87-103
    ([-4, 4, 1], [0, 0, -1], [0, 1, 0], 160.0, 1.0),              # camera.rs:105-122
    ([13, 2, 3], [0, 0, 0], [0, 1, 0], 20.0, 16.0 / 9.0),         # cover camera, 16:9 configs
    ([-2, 0.5, 1], [0, 0, -1], [0, 1, 0], 50.0, 4.0 / 3.0),       # test_scene camera
])
def test_host_camera_equals_oracle_bitwise(args):
    cam = R.camera_from_params(*args)
    p = R.rt_camera_params(R.vec3(args[0]), R.vec3(args[1]), R.vec3(args[2]), args[3], args[4])
    ref = R.rt_camera()
    O.lib().oracle_camera_new(C.byref(p), C.byref(ref))
    assert bytes(cam) == bytes(ref)


def test_shard_rows_partition():
    for h in (1, 2, 7, 600, 1080):
        for world in (1, 2, 3, 4, 8):
            for band in (1, 2, 16):
                rows = [R.shard_rows(h, r, world, band) for r in range(world)]
                assert sum(rows) == h
                idx = np.concatenate([R.shard_row_indices(h, r, world, band) for r in range(world)])
                assert sorted(idx.tolist()) == list(range(h))
                for r in range(world):
                    assert len(R.shard_row_indices(h, r, world, band)) == rows[r]


def test_scene_fixtures_parse_like_the_reference():
    cfg = scenes.cover_config()      # config.rs:249-255 asserts 800x600 on test_scene; cover_scene.json:2-5 ships 64 spp
    assert (cfg["width"], cfg["height"], cfg["samples_per_pixel"], cfg["max_depth"]) == (800, 600, 64, 50)
    kinds = [next(iter(o["material"])) for o in cfg["objects"]]
    assert len(kinds) == 484 and kinds.count("Lambertian") == 407 and kinds.count("Metal") == 56 and kinds.count("Glass") == 21
    t = scenes.test_scene_config()
    assert (t["width"], t["height"]) == (800, 600) and len(t["objects"]) == 7
    sc = scenes.scene("C2")
    assert (sc.c.width, sc.c.height, sc.c.samples_per_pixel, sc.c.max_depth, sc.n_spheres) == (800, 600, 128, 50, 484)
    sc = scenes.scene("C3")
    assert abs(sc.camera_params["aspect"] - 16.0 / 9.0) < 1e-15


def test_rtiow_generator_is_seeded_and_sized():
    a = scenes.rtiow_config(11); b = scenes.rtiow_config(11)
    assert a == b and 470 <= len(a["objects"]) <= 488
    c = scenes.rtiow_config(50)
    assert 9990 <= len(c["objects"]) <= 10004
    assert c["objects"][0]["radius"] == 1000.0 and c["objects"][-1]["material"]["Metal"]["fuzz"] == 0.0


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_render_without_gpu_fails_loudly():
    sc = scenes.cover_scene(16, 12, 1)
    with pytest.raises(R.RtError) as e:
        R.render_rgb8(sc)
    assert e.value.code in (-2, -3) and len(str(e.value)) > 20
    with pytest.raises(R.RtError):
        R.ResidentScene(sc)


def test_invalid_scenes_are_rejected_before_touching_the_device():
    sc = scenes.cover_scene(16, 12, 1)
    sc.c.width = 1            # u = (x+xi)/(w-1): raytracer.rs:199
    with pytest.raises(R.RtError) as e:
        R.render_rgb8(sc)
    assert e.value.code == -1
    sc = scenes.cover_scene(16, 12, 1)
    sc.c.samples_per_pixel = 0
    with pytest.raises(R.RtError) as e:
        R.render_rgb8(sc)
    assert e.value.code == -1
    with pytest.raises(R.RtError) as e:
        R.render_rgb8(scenes.cover_scene(16, 12, 1), R.make_options(rank=2, world=2))
    assert e.value.code == -1


def test_multi_gpu_entry_point_validates_before_touching_the_device():
    sc = scenes.cover_scene(16, 12, 1)
    with pytest.raises(R.RtError) as e:
        R.render_rgb8_multi(sc, 2, R.make_options(rank=1, world=2))     # the call shards the frame itself
    assert e.value.code == -1
    sc.c.samples_per_pixel = 0
    with pytest.raises(R.RtError) as e:
        R.render_rgb8_multi(sc, 0)
    assert e.value.code == -1


def test_texture_buffer_size_is_checked():
    """rt_image.bytes (ABI 2): the callee reads width*height*3 bytes, so a smaller buffer is rejected (the reference would
    panic on the first out-of-bounds texel instead)."""
    img = np.zeros((4, 4, 3), np.uint8)
    cfg = scenes._variant(scenes.cover_config(), 16, 12, 1, 2)
    sc = R.Scene.from_config(cfg)
    sc.c.sky.mode = R.RT_SKY_TEXTURE
    sc.c.sky.tex = R.rt_image(img.ctypes.data, 8, 8, img.size)          # claims 8x8, holds 4x4
    with pytest.raises(R.RtError) as e:
        R.render_rgb8(sc)
    assert e.value.code == -1 and "bytes" in str(e.value)
