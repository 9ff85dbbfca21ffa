"""Host logic of the closest-hit stage (no GPU): structure of the 8-wide BVH rtb200_scene_upload stages, and a float32
emulation of the kernel's conservative tests (slab test per node, 7-FMA sphere test per leaf) that checks the soundness
claim of DESIGN.md §4.2 on random rays: every sphere the exact float64 Sphere::hit accepts is reached by the traversal,
also for scenes far from the origin. (The -m gpu tests assert the end result bit for bit; this one localises a failure.)"""
import numpy as np
import pytest

import rtb200 as R
from rtb200 import scenes
from synth import base_config, mixed_config, _v

f32 = np.float32
EMPTY, LEAF = 0xFFFFFFFF, 0x80000000
U = 2.0 ** -24


def _spheres(sc):
    c = np.array([[s.center.x, s.center.y, s.center.z] for s in sc._spheres[: sc.n_spheres]], np.float64).reshape(-1, 3)
    r = np.array([s.radius for s in sc._spheres[: sc.n_spheres]], np.float64)
    return c, r


def _fma(a, b, c):       # one rounding, like FFMA (the f64 product of two f32 is exact)
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def _leaf_members(b, leaf):
    ids = b["leaf_id"][leaf]
    return ids[ids != EMPTY]


def _subtree_spheres(b, node, out):
    for ref in b["child"][node]:
        if ref == EMPTY:
            continue
        if ref & LEAF:
            out.extend(_leaf_members(b, int(ref & 0x7FFFFFFF)).tolist())
        else:
            _subtree_spheres(b, int(ref), out)


SCENES = [
    lambda: scenes.cover_scene(64, 48, 1),
    lambda: R.Scene.from_config(mixed_config(32, 24, 1, 4, seed=3, n=90)),
    lambda: R.Scene.from_config(scenes._variant(scenes.rtiow_config(20), 32, 24, 1, 4)),
    lambda: R.Scene.from_config(mixed_config(32, 24, 1, 4, seed=4, n=40, offset=(-3.0e6, 1.0e3, 7.0e6))),
    lambda: R.Scene.from_config(mixed_config(16, 12, 1, 2, seed=1, n=5)),
]


@pytest.mark.parametrize("mk", SCENES)
def test_bvh_structure(mk):
    sc = mk()
    b = R.bvh_records(sc)
    c, r = _spheres(sc)
    n = sc.n_spheres
    ids = b["leaf_id"].ravel()
    used = np.concatenate([ids[ids != EMPTY], b["always"]])
    assert sorted(used.tolist()) == list(range(n))                       # every sphere in exactly one leaf slot (or the always-list)
    assert 1 <= b["depth"] <= 32 and b["n_nodes"] >= 1
    g = b["recentre"]
    seen_nodes, seen_leaves = {0}, set()
    for node in range(b["n_nodes"]):
        for k, ref in enumerate(b["child"][node]):
            lo, hi = b["lo"][node][:, k].astype(np.float64), b["hi"][node][:, k].astype(np.float64)
            if ref == EMPTY:
                assert np.all(lo == np.inf) and np.all(hi == -np.inf)   # empty slot: never hit
                continue
            mem = []
            if ref & LEAF:
                leaf = int(ref & 0x7FFFFFFF); assert leaf < b["n_leaves"] and leaf not in seen_leaves; seen_leaves.add(leaf)
                mem = _leaf_members(b, leaf).tolist()
                assert len(mem) >= 1
            else:
                assert node < int(ref) < b["n_nodes"] and int(ref) not in seen_nodes; seen_nodes.add(int(ref))
                _subtree_spheres(b, int(ref), mem)
            cm, rm = c[mem] - g, np.abs(r[mem])
            assert np.all(lo <= (cm - rm[:, None]).min(axis=0)) and np.all(hi >= (cm + rm[:, None]).max(axis=0))   # the child's box holds its spheres
            # ... with the rounding margin of the slab test on the box's side: 32u * max|coordinate|
            bmax = max(np.abs((cm - rm[:, None])).max(), np.abs((cm + rm[:, None])).max())
            assert np.all((cm - rm[:, None]).min(axis=0) - lo >= 31 * U * bmax) and np.all(hi - (cm + rm[:, None]).max(axis=0) >= 31 * U * bmax)
    assert seen_nodes == set(range(b["n_nodes"])) and seen_leaves == set(range(b["n_leaves"]))
    # padding slots of the leaves never hit; the flat records list every sphere in order
    rec = b["leaf_rec"]
    nk = np.stack([rec[:, :, 1, 2], rec[:, :, 1, 3]], axis=2).reshape(b["n_leaves"], -1)
    assert np.all(nk[b["leaf_id"] == EMPTY] == -np.inf)
    flat_nk = np.stack([b["flat"][:, 1, 2], b["flat"][:, 1, 3]], axis=1).ravel()
    assert np.all(np.isfinite(flat_nk[:n])) and np.all(flat_nk[n:] == -np.inf)


def _exact_hits(c, r, o, d):
    """Spheres the reference's f64 test accepts for the ray (sphere.rs:46-58 with t_min 0.001, t_max inf)."""
    oc = o - c
    a = d @ d
    hb = oc @ d
    cc = (oc * oc).sum(axis=1) - r * r
    disc = hb * hb - a * cc
    ok = disc >= 0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    r1, r2 = (-hb - sq) / a, (-hb + sq) / a
    return np.nonzero(ok & ((r1 > 0.001) | (r2 > 0.001)))[0]


def _traverse(b, o, d):
    """The kernel's traversal in emulated float32 (rtb200_wavefront.cu, closest-hit stage). Returns the candidate spheres."""
    g = b["recentre"]
    of = (o - g).astype(f32)
    df = d.astype(f32)
    s = _fma(df[0], df[0], _fma(df[1], df[1], f32(df[2] * df[2])))
    oo = _fma(of[0], of[0], _fma(of[1], of[1], f32(of[2] * of[2])))
    assert 1e-30 < s < 1e30 and oo < 1e30
    dn = (df * f32(1.0 / np.sqrt(np.float64(s)))).astype(f32)          # rsqrtf is within 2 ulp of this
    nod = f32(-_fma(of[0], dn[0], _fma(of[1], dn[1], f32(of[2] * dn[2]))))
    thr = f32(np.nextafter(f32(oo * f32(1.0 - 96.0 * U)), f32(-np.inf)))   # __fmul_rd
    ax = np.where(np.abs(dn) < f32(1e-20), np.copysign(f32(1e-20), dn), dn).astype(f32)
    inv = (f32(1.0) / ax).astype(f32)
    mray = f32(np.nextafter(f32(f32(1.9073486328125e-6) * f32(np.nextafter(np.sqrt(oo, dtype=f32), f32(np.inf)))), f32(np.inf)))
    sm = np.copysign(mray, inv).astype(f32)
    cn = ((of + sm).astype(f32) * (-inv)).astype(f32)
    cf = ((of - sm).astype(f32) * (-inv)).astype(f32)
    neg = np.signbit(inv)
    cands, stack, visited = [], [0], 0
    while stack:
        node = stack.pop(); visited += 1
        lo, hi = b["lo"][node], b["hi"][node]
        near = np.where(neg[:, None], hi, lo); far = np.where(neg[:, None], lo, hi)
        tn = np.stack([_fma(near[a], inv[a], cn[a]) for a in range(3)]).max(axis=0)
        tf = np.stack([_fma(far[a], inv[a], cf[a]) for a in range(3)]).min(axis=0)
        hit = np.maximum(tn, f32(0)) <= tf
        for k in np.nonzero(hit)[0]:
            ref = int(b["child"][node][k])
            assert ref != EMPTY
            if ref & LEAF:
                leaf = ref & 0x7FFFFFFF
                rec = b["leaf_rec"][leaf]                               # [pairs, 2, 4]
                cx = np.stack([rec[:, 0, 0], rec[:, 0, 1]], 1).ravel(); cy = np.stack([rec[:, 0, 2], rec[:, 0, 3]], 1).ravel()
                cz = np.stack([rec[:, 1, 0], rec[:, 1, 1]], 1).ravel(); nk = np.stack([rec[:, 1, 2], rec[:, 1, 3]], 1).ravel()
                bb = _fma(cx, dn[0], _fma(cy, dn[1], _fma(cz, dn[2], nod)))
                tt = _fma(cx, f32(2) * of[0], _fma(cy, f32(2) * of[1], _fma(cz, f32(2) * of[2], nk)))
                with np.errstate(invalid="ignore", over="ignore"):
                    D = _fma(bb, bb, tt)
                ids = b["leaf_id"][leaf]
                cands.extend(ids[(D >= thr) & (ids != EMPTY)].tolist())
            else:
                stack.append(ref)
    return set(cands) | set(b["always"].tolist()), visited


@pytest.mark.parametrize("mk", SCENES[:4])
def test_emulated_traversal_never_drops_a_sphere_the_exact_test_accepts(mk):
    sc = mk()
    b = R.bvh_records(sc)
    c, r = _spheres(sc)
    rng = np.random.default_rng(11)
    cam = np.array([sc.c.camera.origin.x, sc.c.camera.origin.y, sc.c.camera.origin.z])
    total_exact = total_cand = total_nodes = 0
    for i in range(400):
        j = int(rng.integers(len(r)))
        if i % 4 == 0:
            o = cam
            d = (c[j] + rng.normal(size=3) * abs(r[j]) * 0.7) - o            # primary-like ray towards a sphere
        else:
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            o = c[j] + nrm * abs(r[j])                                      # scattered ray leaving a surface point
            d = nrm + rng.normal(size=3) * 0.8
            if i % 8 == 1:
                d = d * np.array([1.0, 1e-9, 1.0])                          # grazing
            if i % 16 == 3:
                d = np.array([0.0, 0.0, 1.0]) * (1 if i % 32 == 3 else -1)  # axis-parallel: zero components
        d = d * float(rng.uniform(0.2, 5.0))                                # directions are not normalised (camera.rs:79-84)
        exact = _exact_hits(c, r, o, d)
        cand, visited = _traverse(b, o, d)
        missing = set(exact.tolist()) - cand
        assert not missing, (i, sorted(missing))
        total_exact += len(exact); total_cand += len(cand); total_nodes += visited
    assert total_exact > 200
    assert total_cand <= 6 * total_exact + 400 * 3                          # the tests prune: few candidates beyond the true hits
    assert total_nodes <= 400 * (3 + 2 * b["depth"]) * 4


def test_degenerate_inputs_build():
    """Coincident spheres, zero / negative radii, a non-finite sphere (always-list), a single sphere, an empty scene."""
    objs = [{"center": _v(0, 0, 0), "radius": 0.5, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}} for _ in range(40)]
    objs += [{"center": _v(1, 0, 0), "radius": 0.0, "material": {"Glass": {"index_of_refraction": 1.5}}},
             {"center": _v(2, 0, 0), "radius": -0.4, "material": {"Glass": {"index_of_refraction": 1.5}}},
             {"center": _v(float("inf"), 0, 0), "radius": 1.0, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}},
             {"center": _v(1e20, 0, 0), "radius": 1.0, "material": {"Lambertian": {"albedo": [0.5, 0.5, 0.5]}}}]
    b = R.bvh_records(R.Scene.from_config(base_config(8, 6, 1, 2, objs)))
    assert sorted(b["always"].tolist()) == [42, 43] and b["depth"] <= 32
    ids = b["leaf_id"].ravel()
    assert sorted(ids[ids != EMPTY].tolist()) == list(range(42))
    one = R.bvh_records(R.Scene.from_config(base_config(8, 6, 1, 2, objs[:1])))
    assert one["n_nodes"] == 1 and one["n_leaves"] == 1 and one["depth"] == 1
    none = R.bvh_records(R.Scene.from_config(base_config(8, 6, 1, 2, [])))
    assert none["n_nodes"] == 0 and none["n_leaves"] == 0


def test_breadth_first_collapse_is_sound_and_bounds_the_depth(monkeypatch):
    """Below wide level 15 the builder expands every inner child twice (three binary levels per wide level), which bounds the
    wide depth by 21 for any input; real scenes never get there, so the test hook RTB200_BVH_AREA_LEVELS=1 forces that collapse
    from the root: the tree must still hold every sphere once, bound its members, and the emulated traversal must still reach
    every sphere the exact test accepts."""
    monkeypatch.setenv("RTB200_BVH_AREA_LEVELS", "1")
    sc = R.Scene.from_config(scenes._variant(scenes.rtiow_config(20), 32, 24, 1, 4))
    b = R.bvh_records(sc)
    monkeypatch.delenv("RTB200_BVH_AREA_LEVELS")
    ref = R.bvh_records(sc)
    assert b["depth"] <= 21 and b["n_leaves"] == ref["n_leaves"] and b["n_nodes"] != ref["n_nodes"]
    ids = b["leaf_id"].ravel()
    assert sorted(ids[ids != EMPTY].tolist()) == list(range(sc.n_spheres))
    c, r = _spheres(sc)
    rng = np.random.default_rng(3)
    cam = np.array([sc.c.camera.origin.x, sc.c.camera.origin.y, sc.c.camera.origin.z])
    n_exact = 0
    for i in range(150):
        j = int(rng.integers(len(r)))
        o = cam if i % 3 == 0 else c[j] + np.array([0.0, abs(r[j]), 0.0])
        d = (c[int(rng.integers(len(r)))] + rng.normal(size=3) * 0.3) - o
        exact = _exact_hits(c, r, o, d)
        cand, _ = _traverse(b, o, d)
        assert not (set(exact.tolist()) - cand)
        n_exact += len(exact)
    assert n_exact > 100
