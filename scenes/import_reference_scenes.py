#!/usr/bin/env python3
"""Imports the reference's INPUT DATA (not source) into scenes/: the two JSON scenes, re-serialised compactly and
gzip-compressed, and the three JPEG textures. Run in the build container where /root/reference exists; the
GPU box only ever reads the committed copies. Also emits the derived BASELINE configs' scene variants lazily
(see scenes/__init__ helpers in rtb200.scenes)."""
import gzip, json, os, shutil, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/raytracer/data"
HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("cover_scene.json", "test_scene.json"):
    with open(os.path.join(REF, name), "rb") as f:
        cfg = json.loads(f.read())
    blob = json.dumps(cfg, separators=(",", ":")).encode()
    with gzip.GzipFile(os.path.join(HERE, name + ".gz"), "wb", mtime=0) as g:
        g.write(blob)
    print(name, len(cfg["objects"]), "objects ->", name + ".gz")
for name in ("earth.jpg", "moon.jpg", "beach.jpg"):
    shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, "data", name))
    os.chmod(os.path.join(HERE, "data", name), 0o644)
    print("copied", name)
