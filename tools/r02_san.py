#!/usr/bin/env python3
"""Small renders for compute-sanitizer (memcheck / racecheck): cover 64x48x4 and a 2-light mixed scene, checked against the oracle."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'rust-raytracer_b200')); sys.path.insert(0, os.path.join(REPO, 'oracle')); sys.path.insert(0, os.path.join(REPO, 'tests'))
import numpy as np
import rtb200 as R
from rtb200 import scenes
import oracle_py as O
from synth import mixed_config, _v

def check(name, sc):
    lin_o, img_o, st_o = O.render(sc)
    lin_g, st_g = R.render_linear(sc)
    img_g, _ = R.render_rgb8(sc)
    ok = np.array_equal(lin_g, lin_o) and np.array_equal(img_g, img_o) and st_g["rays"] == st_o["rays"]
    print(f"{name}: rays {st_g['rays']} exact={ok}", flush=True)
    assert ok

check("cover 64x48x4", scenes.cover_scene(64, 48, 4))
cfg = mixed_config(48, 36, 3, 6, seed=23, n=30)
for k, pos in enumerate([(0.0, 6.0, 0.0), (-4.0, 3.0, 5.0)]):
    cfg["objects"].insert(3 + 5 * k, {"center": _v(*pos), "radius": 1.0 + 0.5 * k, "material": {"Light": {}}})
check("2 lights 48x36x3", R.Scene.from_config(cfg))
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cfg = scenes._variant(scenes.rtiow_config(50), 64, 36, 2, 50)
    check("rtiow 10k 64x36x2", R.Scene.from_config(cfg))
