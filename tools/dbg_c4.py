import sys, os
sys.path.insert(0, 'rust-raytracer_b200'); sys.path.insert(0, 'tests')
import rtb200 as R
from rtb200 import scenes
sc4 = R.Scene.from_config(scenes._variant(scenes.rtiow_config(50), 128, 72, 3, 50))
def c4(tag):
    lin, st = R.render_linear(sc4)
    print(tag, 'c4 rays',st['rays'],'cand/ray',round(st['candidates']/st['rays'],3),'clus/ray',round(st['clusters']/st['rays'],2), flush=True)
c4('fresh')
R.render_linear(scenes.cover_scene(200,150,8)); c4('after cover 1'); c4('after cover 2'); c4('after cover 3')
rs = R.ResidentScene(sc4)
import torch
out = torch.empty(72*128*3, dtype=torch.uint8, device='cuda')
for i in range(3):
    st = rs.render(out.data_ptr()); print('resident', i, round(st['candidates']/st['rays'],3), flush=True)
R.render_linear(scenes.cover_scene(200,150,8))
for i in range(2):
    st = rs.render(out.data_ptr()); print('resident after cover', i, round(st['candidates']/st['rays'],3), flush=True)
