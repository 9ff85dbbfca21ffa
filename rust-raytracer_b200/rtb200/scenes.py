"""Concrete inputs of the BASELINE.json configs built from the committed scene fixtures.

C1  test_scene   400x300   16 spp depth 8   (CPU plumbing case of the reference)
C2  cover_scene  800x600  128 spp depth 50  (the config the headline metric is quoted on)
C3  cover_scene 1920x1080 512 spp depth 50
C4  RTIOW-10k   1920x1080 1024 spp depth 50 (seeded restatement of config.rs:149-226 on a [-50,50)^2 grid)
C5  cover_scene 3840x2160 4096 spp depth 50
"""
from __future__ import annotations

import copy
import os

import numpy as np

from . import Scene, read_config

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SCENES_DIR = os.path.join(REPO, "scenes")


def cover_config() -> dict:
    return read_config(os.path.join(SCENES_DIR, "cover_scene.json.gz"))


def test_scene_config() -> dict:
    return read_config(os.path.join(SCENES_DIR, "test_scene.json.gz"))


def _variant(cfg: dict, w: int, h: int, spp: int, depth: int) -> dict:
    cfg = copy.deepcopy(cfg)
    cfg["width"], cfg["height"], cfg["samples_per_pixel"], cfg["max_depth"] = w, h, spp, depth
    cfg["camera"]["aspect"] = float(w) / float(h)   # aspect is a camera field independent of width/height (camera.rs:26,54)
    return cfg


def rtiow_config(half: int = 50, seed: int = 20240924) -> dict:
    """Seeded restatement of `_make_cover_world` (reference config.rs:149-226) on a [-half, half)^2 grid.
    half=11 reproduces the reference's grid size (484 spheres); half=50 gives the ~10k-sphere BASELINE scene."""
    rng = np.random.Generator(np.random.Philox(seed))
    objs = [{"center": {"x": 0.0, "y": -1000.0, "z": 0.0}, "radius": 1000.0, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}}]
    for a in range(-half, half):
        for b in range(-half, half):
            choose = float(rng.random())
            cx = a + 0.9 * float(rng.random()); cz = b + 0.9 * float(rng.random())
            if ((cx - 4.0) ** 2 + (0.2 - 0.2) ** 2 + cz ** 2) ** 0.5 < 0.9:
                continue
            c = {"x": cx, "y": 0.2, "z": cz}
            f32 = lambda v: float(np.float32(v))
            if choose < 0.8:
                alb = [f32(np.float32(rng.random(dtype=np.float32)) * np.float32(rng.random(dtype=np.float32))) for _ in range(3)]
                m = {"Lambertian": {"albedo": alb}}
            elif choose < 0.95:
                alb = [f32(np.float32(0.5) * (np.float32(1.0) + np.float32(rng.random(dtype=np.float32)))) for _ in range(3)]
                m = {"Metal": {"albedo": alb, "fuzz": 0.5 * float(rng.random())}}
            else:
                m = {"Glass": {"index_of_refraction": 1.5}}
            objs.append({"center": c, "radius": 0.2, "material": m})
    objs.append({"center": {"x": 0.0, "y": 1.0, "z": 0.0}, "radius": 1.0, "material": {"Glass": {"index_of_refraction": 1.5}}})
    objs.append({"center": {"x": -4.0, "y": 1.0, "z": 0.0}, "radius": 1.0, "material": {"Lambertian": {"albedo": [0.4, 0.2, 0.1]}}})
    objs.append({"center": {"x": 4.0, "y": 1.0, "z": 0.0}, "radius": 1.0, "material": {"Metal": {"albedo": [0.7, 0.6, 0.5], "fuzz": 0.0}}})
    return {"width": 800, "height": 600, "samples_per_pixel": 64, "max_depth": 50, "sky": {"texture": ""},
            "camera": {"look_from": {"x": 13.0, "y": 2.0, "z": 3.0}, "look_at": {"x": 0.0, "y": 0.0, "z": 0.0},
                       "vup": {"x": 0.0, "y": 1.0, "z": 0.0}, "vfov": 20.0, "aspect": 800.0 / 600.0},
            "objects": objs}


def config(name: str) -> dict:
    name = name.upper()
    if name == "C1":
        return _variant(test_scene_config(), 400, 300, 16, 8)
    if name == "C2":
        return _variant(cover_config(), 800, 600, 128, 50)
    if name == "C3":
        return _variant(cover_config(), 1920, 1080, 512, 50)
    if name == "C4":
        return _variant(rtiow_config(50), 1920, 1080, 1024, 50)
    if name == "C5":
        return _variant(cover_config(), 3840, 2160, 4096, 50)
    # reduced sizes of the big configs: golden-frame parity cases (tests/golden/frames.json) and profiling targets
    if name == "C3S":
        return _variant(cover_config(), 1920, 1080, 4, 50)
    if name == "C4S":
        return _variant(rtiow_config(50), 480, 270, 4, 50)
    if name == "C4M":
        return _variant(rtiow_config(50), 960, 540, 16, 50)
    if name == "C5S":
        return _variant(cover_config(), 3840, 2160, 1, 50)
    raise KeyError(name)


def scene(name: str) -> Scene:
    return Scene.from_config(config(name), SCENES_DIR)


def cover_scene(w: int, h: int, spp: int, depth: int = 50) -> Scene:
    return Scene.from_config(_variant(cover_config(), w, h, spp, depth), SCENES_DIR)
