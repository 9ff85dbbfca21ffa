#!/usr/bin/env python3
"""Render a BASELINE config resident N times (profiling / timing target). usage: render_once.py [C2] [reps] [variant]"""
import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'rust-raytracer_b200'))
import rtb200 as R
from rtb200 import scenes
import torch
name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
world = int(sys.argv[4]) if len(sys.argv) > 4 else 1      # render shard 0 of `world` (what one GPU of an N-GPU run does)
if 'x' in name:
    w, h, spp = (int(v) for v in name.split('x'))
    sc = scenes.cover_scene(w, h, spp)
else:
    sc = scenes.scene(name)
rs = R.ResidentScene(sc, R.make_options(variant=variant, rank=0, world=world))
ki = rs.kernel_info()
print(f"{name}: {ki['name']} regs={ki['registers']} local={ki['local_bytes']} smem={ki['smem_bytes']} grid={ki['grid']} ({ki['ctas_per_sm']}/SM) smem_mask={ki['smem_mask']} "
      f"bvh nodes={ki['bvh_nodes']} leaves={ki['bvh_leaves']} depth={ki['bvh_depth']}", flush=True)
out = torch.empty(sc.c.height * sc.c.width * 3, dtype=torch.uint8, device='cuda')
if world > 1:
    name = f"{name}/shard0of{world}"
for i in range(reps):
    st = rs.render(out.data_ptr())
    r = max(st['rays'], 1)
    print(f"{name}: rays={st['rays']} device_ms={st['device_ms']:.3f} trace_ms={st['trace_ms']:.3f} Mrays/s={st['rays']/st['device_ms']/1e3:.1f} "
          f"f64tests/ray={st['candidates']/r:.2f} leaves/ray={st['clusters']/r:.2f} nodes/ray={st['nodes']/r:.2f}", flush=True)
