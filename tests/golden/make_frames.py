#!/usr/bin/env python3
"""Writes tests/golden/frames.json: SHA-256 of the ORACLE's linear-f32 and RGB8 frames and its ray count for the
BASELINE configs at sizes the CPU oracle can finish (full-size C2; C3/C4/C5 at reduced samples or size, same scene,
aspect and depth). ORACLE-generated regression pins (the reference ships no golden image and cannot be built here);
asserted by the `-m gpu` tests and by bench.py on rank 0 at every N ("golden": "match").
Usage: python tests/golden/make_frames.py [names...]      (all: about two minutes on 8 cores)
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "rust-raytracer_b200")); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np          # noqa: E402
import oracle_py as O       # noqa: E402
from rtb200 import scenes   # noqa: E402

NAMES = ["C1", "C2", "C3S", "C4S", "C5S"]
OUT = os.path.join(HERE, "frames.json")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


if __name__ == "__main__":
    names = [n.upper() for n in sys.argv[1:]] or NAMES
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        cfg = scenes.config(name)
        sc = scenes.scene(name)
        t = time.time()
        lin, img, st = O.render(sc)
        db[name] = {
            "workload": f"{len(cfg['objects'])} spheres {cfg['width']}x{cfg['height']} {cfg['samples_per_pixel']}spp depth {cfg['max_depth']} seed {sc.seed:#x}",
            "width": cfg["width"], "height": cfg["height"], "samples_per_pixel": cfg["samples_per_pixel"], "max_depth": cfg["max_depth"],
            "n_spheres": len(cfg["objects"]), "seed": int(sc.seed),
            "rays": int(st["rays"]), "samples": int(st["samples"]),
            "sha256_rgb8": sha(img), "sha256_linear_f32": sha(lin), "made_by": "oracle/liboracle.so (tests/golden/make_frames.py)",
        }
        print(name, db[name]["workload"], "rays", st["rays"], f"{time.time() - t:.1f}s", flush=True)
        with open(OUT, "w") as f:
            json.dump(db, f, indent=1, sort_keys=True)
