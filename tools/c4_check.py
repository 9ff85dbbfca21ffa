#!/usr/bin/env python3
"""C4 (RTIOW 10k spheres) at reduced size: parity vs the oracle and timing (run under gpurun with timeout)."""
import sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'rust-raytracer_b200')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import numpy as np
import rtb200 as R
from rtb200 import scenes
import oracle_py as O
cfg = scenes._variant(scenes.rtiow_config(50), 160, 90, 4, 50)
sc = R.Scene.from_config(cfg)
print("spheres", sc.n_spheres, flush=True)
lin_g, st = R.render_linear(sc)
print(f"GPU 160x90x4: rays={st['rays']} trace_ms={st['trace_ms']:.2f} Mrays/s={st['rays']/st['trace_ms']/1e3:.1f} cand/ray={st['candidates']/st['rays']:.2f} clusters/ray={st['clusters']/st['rays']:.2f}", flush=True)
t = time.time(); lin_o, _, so = O.render(sc); print(f"oracle {time.time()-t:.1f}s rays={so['rays']} {so['rays']/so['render_ms']/1e3:.3f} Mrays/s", flush=True)
print("max|d|", float(np.abs(lin_g - lin_o).max()), "rays equal", st['rays'] == so['rays'], flush=True)
cfg = scenes._variant(scenes.rtiow_config(50), 960, 540, 16, 50)
sc = R.Scene.from_config(cfg)
rs = R.ResidentScene(sc)
import torch
out = torch.empty(sc.c.height * sc.c.width * 3, dtype=torch.uint8, device='cuda')
for i in range(2):
    st = rs.render(out.data_ptr())
    print(f"GPU 960x540x16: rays={st['rays']} trace_ms={st['trace_ms']:.2f} Mrays/s={st['rays']/st['trace_ms']/1e3:.1f} cand/ray={st['candidates']/st['rays']:.2f} clusters/ray={st['clusters']/st['rays']:.2f}", flush=True)
