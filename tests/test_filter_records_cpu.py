"""Host logic of the closest-hit filter (no GPU): structure of the two-level records and a float32 emulation of the
7-FMA test that checks the soundness claim of DESIGN.md §4.2 — whenever the exact (float64) ray/sphere discriminant is
non-negative, the f32 test accepts the sphere AND the bound of its cluster, also far from the origin."""
import numpy as np
import pytest

import rtb200 as R
from rtb200 import scenes
from synth import mixed_config


def _unpack(rec):       # [pairs, 2, 4] -> per-record (cx, cy, cz, nk), record order
    a, b = rec[:, 0, :], rec[:, 1, :]
    cx = np.stack([a[:, 0], a[:, 1]], 1).ravel(); cy = np.stack([a[:, 2], a[:, 3]], 1).ravel()
    cz = np.stack([b[:, 0], b[:, 1]], 1).ravel(); nk = np.stack([b[:, 2], b[:, 3]], 1).ravel()
    return cx, cy, cz, nk


def _spheres(sc):
    c = np.array([[s.center.x, s.center.y, s.center.z] for s in sc._spheres[: sc.n_spheres]], np.float64)
    r = np.array([s.radius for s in sc._spheres[: sc.n_spheres]], np.float64)
    return c, r


@pytest.mark.parametrize("mk", [lambda: scenes.cover_scene(64, 48, 1), lambda: R.Scene.from_config(mixed_config(32, 24, 1, 4, seed=3, n=90)),
                               lambda: R.Scene.from_config(scenes._variant(scenes.rtiow_config(20), 32, 24, 1, 4))])
def test_two_level_structure(mk):
    sc = mk()
    fr = R.filter_records(sc)
    c, r = _spheres(sc)
    assert fr["two_level"] and fr["cluster_size"] == 4 and fr["n_clusters"] % 4 == 0 and fr["n_pairs"] % 8 == 0
    slots = fr["slot_to_sphere"]
    used = slots[slots != 0xFFFF]
    assert sorted(used.tolist()) == list(range(sc.n_spheres))          # every sphere in exactly one slot
    g = fr["recentre"]
    cx, cy, cz, nk = _unpack(fr["first"])
    for k in range(fr["n_clusters"]):
        mem = slots[k][slots[k] != 0xFFFF]
        if len(mem) == 0:
            assert nk[k] == -np.inf                                     # padding cluster never hits
            continue
        cc = np.array([cx[k], cy[k], cz[k]], np.float64) + g
        K = -nk[k]                                                      # ~ |c'|^2 - R^2 (minus the safety term)
        R2 = float(cx[k]) ** 2 + float(cy[k]) ** 2 + float(cz[k]) ** 2 - K
        assert R2 > 0
        dist = np.linalg.norm(c[mem] - cc, axis=1) + np.abs(r[mem])
        assert np.all(dist <= np.sqrt(R2) * (1 + 1e-5) + 1e-4), (k, dist.max(), np.sqrt(R2))   # members inside the bound
    # never-hit padding of the second level
    s2 = fr["second"].reshape(-1, 2, 4)
    _, _, _, nk2 = _unpack(s2)
    assert np.all(nk2[(slots.ravel() == 0xFFFF)] == -np.inf)
    assert not R.filter_records(sc, R.RT_VARIANT_BRUTE_FORCE)["two_level"]
    small = R.Scene.from_config(mixed_config(16, 12, 1, 2, seed=1, n=5))
    assert not R.filter_records(small)["two_level"]                     # n <= 32: plain scan


def _f32_test(cx, cy, cz, nk, o32, d32):
    """Emulates the kernel's test in float32 (products and sums rounded separately: more roundings than the FMA chains)."""
    f = np.float32
    inv = f(1.0) / np.sqrt(d32[0] * d32[0] + d32[1] * d32[1] + d32[2] * d32[2], dtype=f)
    dn = (d32 * inv).astype(f)
    nod = -(o32[0] * dn[0] + o32[1] * dn[1] + o32[2] * dn[2]).astype(f)
    oo = (o32[0] * o32[0] + o32[1] * o32[1] + o32[2] * o32[2]).astype(f)
    thr = np.nextafter(f(oo * f(1.0 - 96.0 * 2.0 ** -24)), f(-np.inf))
    b = (cz * dn[2] + nod).astype(f); b = (cy * dn[1] + b).astype(f); b = (cx * dn[0] + b).astype(f)
    t = (cz * f(2) * o32[2] + nk).astype(f); t = (cy * f(2) * o32[1] + t).astype(f); t = (cx * f(2) * o32[0] + t).astype(f)
    D = (b * b + t).astype(f)
    return D >= thr


@pytest.mark.parametrize("offset", [(0.0, 0.0, 0.0), (3.0e3, -2.0e2, 1.0e3), (2.0e6, 1.0e3, -5.0e6)])
def test_filter_never_rejects_what_the_exact_test_accepts(offset):
    sc = R.Scene.from_config(mixed_config(32, 24, 1, 4, seed=5, n=120, offset=offset))
    fr = R.filter_records(sc)
    c, r = _spheres(sc)
    g = fr["recentre"]
    ccx, ccy, ccz, cnk = _unpack(fr["first"])
    scx, scy, scz, snk = _unpack(fr["second"].reshape(-1, 2, 4))
    slots = fr["slot_to_sphere"].ravel()
    cluster_of = np.full(sc.n_spheres, -1); slot_of = np.full(sc.n_spheres, -1)
    for i, sp in enumerate(slots):
        if sp != 0xFFFF:
            cluster_of[sp] = i // fr["cluster_size"]; slot_of[sp] = i
    rng = np.random.default_rng(11)
    off = np.array(offset)
    accepted = 0
    for it in range(4000):
        if it % 2:
            tgt = c[rng.integers(sc.n_spheres)] + rng.normal(size=3) * 0.6          # aimed near a sphere: many silhouette cases
            o = off + rng.uniform(-14, 14, 3); d = (tgt - o) * rng.uniform(0.05, 3.0)
        else:
            o = off + rng.uniform(-14, 14, 3); d = rng.normal(size=3) * rng.uniform(0.05, 3.0)
        oc = o - c
        a = d @ d; hb = oc @ d; cc = (oc * oc).sum(1) - r * r
        disc = hb * hb - a * cc                                                     # the reference's discriminant, float64
        hits = np.where(disc >= 0)[0]
        if len(hits) == 0:
            continue
        o32 = (o - g).astype(np.float32); d32 = d.astype(np.float32)
        ok_c = _f32_test(ccx, ccy, ccz, cnk, o32, d32)
        ok_s = _f32_test(scx, scy, scz, snk, o32, d32)
        assert np.all(ok_s[slot_of[hits]]), "second level rejected a sphere the exact test accepts"
        assert np.all(ok_c[cluster_of[hits]]), "first level rejected the cluster of a sphere the exact test accepts"
        accepted += len(hits)
    assert accepted > 2000
