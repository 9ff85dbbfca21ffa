"""Two restatements of the reference's render loop, written separately, must produce the same bits.

The C++ oracle (oracle/rt_oracle.hpp) is what every GPU test is compared with; the reference itself cannot run here, so the
oracle is pinned by the reference's known-answer tests - and, in this file, by a second restatement in pure Python
(tests/py_restatement.py: recursive like the reference, no code shared with the oracle). Linear f32 frames, RGB8 frames and
ray counts must be identical on small frames of: the cover scene, a mixed-material scene with lights (shadow-ray recursion,
`depth > max_depth - 2` incl. its usize wrap), a black-sky scene, and the reference's test_scene (textures, sky texture,
light, hollow glass)."""
import numpy as np
import pytest

import oracle_py as O
import rtb200 as R
from rtb200 import scenes
from py_restatement import World
from synth import mixed_config, _v


def _both(cfg, base_dir=None):
    sc = R.Scene.from_config(cfg, base_dir or scenes.SCENES_DIR)
    lin_o, img_o, st_o = O.render(sc)
    tex = {}
    k = 0
    for i, o in enumerate(cfg["objects"]):
        if "Texture" in o["material"]:
            tex[i] = sc._tex_arrays[k]; k += 1
    w = World(cfg, textures=tex, sky_texture=sc._sky_array, seed=sc.seed)
    lin_p, img_p, rays_p = w.render()
    return (lin_o, img_o, st_o["rays"]), (lin_p, img_p, rays_p)


def _assert_same(a, b):
    assert a[2] == b[2], ("ray counts", a[2], b[2])
    assert np.array_equal(a[0], b[0]), f"linear frames differ: max |d| = {np.abs(a[0] - b[0]).max()}"
    assert np.array_equal(a[1], b[1])


def test_cover_scene():
    _assert_same(*_both(scenes._variant(scenes.cover_config(), 24, 18, 3, 50)))


@pytest.mark.parametrize("n_lights,depth,sky", [(1, 5, "gradient"), (2, 4, "none"), (1, 1, "gradient"), (1, 2, "gradient")])
def test_mixed_materials_with_lights(n_lights, depth, sky):
    cfg = mixed_config(10, 8, 3, depth, seed=31, n=10, sky=sky)
    for k, pos in enumerate([(0.0, 6.0, 0.0), (-4.0, 3.0, 5.0)][:n_lights]):
        cfg["objects"].insert(2 + 3 * k, {"center": _v(*pos), "radius": 1.0 + 0.5 * k, "material": {"Light": {}}})
    _assert_same(*_both(cfg))


def test_reference_test_scene_with_textures():
    _assert_same(*_both(scenes._variant(scenes.test_scene_config(), 40, 30, 3, 8)))
