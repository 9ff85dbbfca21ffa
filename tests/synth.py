"""Synthetic scene builders for the parity tests (edge cases the reference's data files do not reach)."""
import numpy as np


def _v(x, y, z):
    return {"x": float(x), "y": float(y), "z": float(z)}


def base_config(w, h, spp, depth, objects, sky="gradient", look_from=(13, 2, 3), look_at=(0, 0, 0), vfov=20.0):
    cfg = {"width": w, "height": h, "samples_per_pixel": spp, "max_depth": depth,
           "camera": {"look_from": _v(*look_from), "look_at": _v(*look_at), "vup": _v(0, 1, 0), "vfov": vfov, "aspect": w / h},
           "objects": objects}
    if sky == "gradient":
        cfg["sky"] = {"texture": ""}
    elif sky == "none":
        cfg["sky"] = None
    return cfg


def mixed_config(w, h, spp, depth, seed=1, n=60, offset=(0.0, 0.0, 0.0), sky="gradient"):
    """Random Lambertian/Metal/Glass spheres incl. a hollow (negative-radius) glass shell, two coincident
    spheres (tie on equal t -> first index must win, raytracer.rs:52-56), a sphere around the camera and a
    fuzz-0 mirror; `offset` translates the whole scene (stresses the f32 filter far from the origin)."""
    rng = np.random.default_rng(seed)
    ox, oy, oz = offset
    objs = [{"center": _v(ox, oy - 1000.0, oz), "radius": 1000.0, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}}]
    for i in range(n):
        c = _v(ox + rng.uniform(-6, 6), oy + rng.uniform(0.15, 1.2), oz + rng.uniform(-6, 6))
        r = float(rng.uniform(0.15, 0.6))
        k = rng.uniform()
        if k < 0.5:
            m = {"Lambertian": {"albedo": [float(np.float32(x)) for x in rng.uniform(0, 1, 3)]}}
        elif k < 0.8:
            m = {"Metal": {"albedo": [float(np.float32(x)) for x in rng.uniform(0.4, 1, 3)], "fuzz": float(rng.uniform(0, 0.6)) if rng.uniform() < 0.8 else 0.0}}
        else:
            m = {"Glass": {"index_of_refraction": float(rng.choice([1.5, 1.33, 2.4]))}}
        objs.append({"center": c, "radius": r, "material": m})
    # hollow glass: outer r, inner -0.9 r (test_scene.json:137 pattern)
    objs.append({"center": _v(ox + 1.5, oy + 1.0, oz + 1.0), "radius": 1.0, "material": {"Glass": {"index_of_refraction": 1.5}}})
    objs.append({"center": _v(ox + 1.5, oy + 1.0, oz + 1.0), "radius": -0.9, "material": {"Glass": {"index_of_refraction": 1.5}}})
    # coincident pair with different materials: first index wins every tie
    objs.append({"center": _v(ox - 2.0, oy + 0.7, oz + 2.0), "radius": 0.7, "material": {"Lambertian": {"albedo": [0.9, 0.1, 0.1]}}})
    objs.append({"center": _v(ox - 2.0, oy + 0.7, oz + 2.0), "radius": 0.7, "material": {"Metal": {"albedo": [0.1, 0.9, 0.1], "fuzz": 0.0}}})
    # camera sits inside a big glass bubble
    objs.append({"center": _v(ox + 13, oy + 2, oz + 3), "radius": 2.0, "material": {"Glass": {"index_of_refraction": 1.1}}})
    return base_config(w, h, spp, depth, objs, sky=sky, look_from=(ox + 13, oy + 2, oz + 3), look_at=(ox, oy, oz))
