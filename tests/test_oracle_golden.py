"""The oracle reproduces the committed image fixtures bit for bit (regression pin of the checker itself), and its
statistics match the path structure SURVEY.md §8d measured on the cover scene."""
import os

import numpy as np
import pytest

import oracle_py as O
import rtb200 as R
from rtb200 import scenes
from synth import mixed_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "cover_40x30_s4": lambda: scenes.cover_scene(40, 30, 4),
    "cover_64x48_s2_d3": lambda: scenes.cover_scene(64, 48, 2, depth=3),
    "mixed_48x36_s3": lambda: R.Scene.from_config(mixed_config(48, 36, 3, 12, seed=11), scenes.SCENES_DIR),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    lin, img, st = O.render(CASES[name]())
    assert st["rays"] == int(g["rays"])
    assert np.array_equal(lin, g["linear"]) and np.array_equal(img, g["rgb8"])


def test_oracle_is_thread_count_invariant_and_row_local():
    sc = scenes.cover_scene(48, 36, 3)
    a, ia, sa = O.render(sc, threads=1)
    b, ib, sb = O.render(sc, threads=4)
    assert np.array_equal(a, b) and np.array_equal(ia, ib) and sa["rays"] == sb["rays"]
    c, _, sc_ = O.render(sc, y0=10, y1=20)
    assert np.array_equal(c[10:20], a[10:20]) and not c[:10].any()


def test_cover_scene_path_statistics():
    sc = scenes.cover_scene(160, 120, 8)
    _, _, st = O.render(sc)
    rho = st["rays"] / st["samples"]
    assert 2.55 < rho < 2.8                        # 2.672 rays/sample
    assert st["term_sky"] / st["samples"] > 0.995  # 99.92 % of paths end in the sky
    assert st["hits"][R.RT_TEXTURE] == 0 and st["hits"][R.RT_LIGHT] == 0
    assert st["draws"] > 2 * st["samples"]


def test_depth_limit_and_black_sky():
    sc = scenes.cover_scene(32, 24, 2, depth=1)
    lin, _, st = O.render(sc)
    assert st["rays"] == st["samples"]             # depth 1: only primary rays
    cfg = scenes._variant(scenes.cover_config(), 32, 24, 2, 5); cfg["sky"] = None
    lin, img, _ = O.render(R.Scene.from_config(cfg))
    assert not lin.any() and not img.any()         # no sky, no lights -> black frame (raytracer.rs:138-140)


@pytest.mark.parametrize("name", ["C1", "C4S"])
def test_oracle_reproduces_the_committed_frame_hashes(name):
    """tests/golden/frames.json (the hashes the GPU tests and bench.py assert) is what the oracle renders today."""
    import hashlib
    import json
    with open(os.path.join(GOLD, "frames.json")) as f:
        g = json.load(f)[name]
    lin, img, st = O.render(scenes.scene(name))
    assert st["rays"] == g["rays"]
    assert hashlib.sha256(img.tobytes()).hexdigest() == g["sha256_rgb8"] and hashlib.sha256(lin.tobytes()).hexdigest() == g["sha256_linear_f32"]
