"""The CPU oracle built with UndefinedBehaviorSanitizer (incl. float-cast-overflow: the reference's `as usize` / `as u8`
casts are DEFINED in Rust — saturating — while the same cast is undefined in C++, so the restatement must never rely on it)
renders the committed golden cases without a report and bit-identically to the committed hashes / fixtures."""
import os
import shutil
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "rust-raytracer_b200")); sys.path.insert(0, os.path.join(REPO, "tests"))
import oracle_py as O
import rtb200 as R
from rtb200 import scenes
from synth import mixed_config, base_config
assert os.environ["RTB200_ORACLE_LIB"] == O.LIB_PATH
gold = os.path.join(REPO, "tests", "golden")
frames = json.load(open(os.path.join(gold, "frames.json")))
# textures + sky texture + atan2 + hollow glass (C1), 10,000 spheres (C4S)
for name in ("C1", "C4S"):
    lin, img, st = O.render(scenes.scene(name))
    g = frames[name]
    assert st["rays"] == g["rays"], name
    assert hashlib.sha256(img.tobytes()).hexdigest() == g["sha256_rgb8"], name
    assert hashlib.sha256(lin.tobytes()).hexdigest() == g["sha256_linear_f32"], name
for name, sc in (("cover_40x30_s4", scenes.cover_scene(40, 30, 4)),
                 ("mixed_48x36_s3", R.Scene.from_config(mixed_config(48, 36, 3, 12, seed=11), scenes.SCENES_DIR))):
    g = np.load(os.path.join(gold, name + ".npz"))
    lin, img, st = O.render(sc)
    assert st["rays"] == int(g["rays"]) and np.array_equal(lin, g["linear"]) and np.array_equal(img, g["rgb8"]), name
# lights (shadow recursion, max_depth 1 and 2 -> the usize wrap of `depth > max_depth - 2`), black sky, degenerate geometry
v = lambda x, y, z: {"x": float(x), "y": float(y), "z": float(z)}
for depth in (1, 2, 6):
    objs = [{"center": v(0, -1000, 0), "radius": 1000.0, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}},
            {"center": v(0, 1, 0), "radius": 1.0, "material": {"Glass": {"index_of_refraction": 1.5}}},
            {"center": v(-3, 1, 1), "radius": 1.0, "material": {"Metal": {"albedo": [0.7, 0.6, 0.5], "fuzz": 0.0}}},
            {"center": v(2, 4, 2), "radius": 0.7, "material": {"Light": {}}},
            {"center": v(-2, 5, -2), "radius": 0.5, "material": {"Light": {}}},
            {"center": v(1, 0.5, 2), "radius": 0.0, "material": {"Lambertian": {"albedo": [0.1, 0.9, 0.1]}}},
            {"center": v(float("nan"), 0, 0), "radius": 1.0, "material": {"Lambertian": {"albedo": [0.1, 0.9, 0.1]}}},
            {"center": v(1e300, 1e300, -1e300), "radius": 1e300, "material": {"Metal": {"albedo": [0.7, 0.6, 0.5], "fuzz": 1.0}}}]
    for sky in ("gradient", "none"):
        lin, img, st = O.render(R.Scene.from_config(base_config(40, 30, 3, depth, objs, sky=sky)))
        assert st["rays"] >= st["samples"]
print("ubsan-oracle: ok")
"""


def test_oracle_under_ubsan(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    so = str(tmp_path / "liboracle_ubsan.so")
    b = subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                        "-fsanitize=undefined,float-cast-overflow", "-fno-sanitize-recover=all", "-shared", "-o", so,
                        os.path.join(REPO, "oracle", "rt_oracle_capi.cpp")], capture_output=True, text=True)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip("UBSan runtime not installed: " + b.stderr[-200:])
    assert b.returncode == 0, b.stderr[-3000:]
    env = dict(os.environ, RTB200_ORACLE_LIB=so, UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", "REPO = %r\n" % REPO + CHILD], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "ubsan-oracle: ok" in r.stdout, (r.stdout[-1500:] + "\n" + r.stderr[-6000:])
    assert "runtime error" not in r.stderr
