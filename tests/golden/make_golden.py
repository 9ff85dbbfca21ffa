#!/usr/bin/env python3
"""Regenerates the image fixtures under tests/golden/ with the CPU oracle (oracle/liboracle.so).

The reference itself cannot run here (no Rust toolchain) and ships no golden image, so these fixtures are
ORACLE-generated regression pins, not reference outputs; the oracle in turn is pinned by reference_kat.json.
Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "rust-raytracer_b200")); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import oracle_py as O   # noqa: E402
import rtb200 as R      # noqa: E402
from rtb200 import scenes  # noqa: E402
from synth import mixed_config  # noqa: E402

CASES = {
    "cover_40x30_s4": lambda: scenes.cover_scene(40, 30, 4),            # the reference's own smoke size (raytracer.rs:277-284)
    "cover_64x48_s2_d3": lambda: scenes.cover_scene(64, 48, 2, depth=3),
    "mixed_48x36_s3": lambda: R.Scene.from_config(mixed_config(48, 36, 3, 12, seed=11), scenes.SCENES_DIR),
}
if __name__ == "__main__":
    for name, mk in CASES.items():
        sc = mk()
        lin, img, st = O.render(sc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), linear=lin, rgb8=img, rays=np.uint64(st["rays"]), seed=np.uint64(sc.seed))
        print(name, lin.shape, "rays", st["rays"])
