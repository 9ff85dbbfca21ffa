#!/usr/bin/env python3
"""Quick GPU-vs-oracle check + timing (development aid; the real tests live in tests/)."""
import sys, os, time, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'rust-raytracer_b200')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import numpy as np
import rtb200 as R
from rtb200 import scenes
import oracle_py as O

def cmp(name, sc, variant=0):
    lin_o, img_o, st_o = O.render(sc)
    lin_g, st_g = R.render_linear(sc, R.make_options(variant=variant))
    img_g, _ = R.render_rgb8(sc, R.make_options(variant=variant))
    d = np.abs(lin_g - lin_o)
    print(f"{name}: variant={variant} max|dlin|={d.max():.3e} n_diff={(d>0).sum()} rays gpu={st_g['rays']} oracle={st_o['rays']} "
          f"u8 maxdiff={np.abs(img_g.astype(int)-img_o.astype(int)).max()} cand/ray={st_g['candidates']/max(st_g['rays'],1):.2f} trace_ms={st_g['trace_ms']:.2f}", flush=True)
    return d.max()

sc = scenes.cover_scene(200, 150, 8)
cmp("cover200x150x8", sc, R.RT_VARIANT_EXACT_F64)
cmp("cover200x150x8", sc, R.RT_VARIANT_FILTERED)
cmp("cover200x150x8", sc, R.RT_VARIANT_BRUTE_FORCE)
sc = scenes.cover_scene(64, 48, 4, depth=3)
cmp("cover64x48 depth3", sc, R.RT_VARIANT_FILTERED)
if len(sys.argv) > 1 and sys.argv[1] == 'time':
    sc = scenes.scene('C2')
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rs = R.ResidentScene(sc, R.make_options(variant=variant))
    import torch
    out = torch.empty(sc.c.height*sc.c.width*3, dtype=torch.uint8, device='cuda')
    for i in range(4):
        st = rs.render(out.data_ptr())
        print(f"C2 resident: rays={st['rays']} device_ms={st['device_ms']:.2f} trace_ms={st['trace_ms']:.2f} Mrays/s={st['rays']/st['device_ms']/1e3:.1f} cand/ray={st['candidates']/st['rays']:.2f}", flush=True)
    img, st = R.render_rgb8(sc)
    print("C2 host e2e:", json.dumps(st))
    R.write_png(os.path.join(REPO, 'gpurun_out', 'c2.png'), img)
