import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quadraturefields_b200 import scene
from quadraturefields_b200.utils import GraphedTrainStep, render_train
dev = torch.device("cuda:0")
sc = scene.make_scene("smoke", device=dev)
rf, mi = sc.radiance_field, sc.mesh_intersect
params = [rf.mlp_base.params, rf.mlp_head.params]
for p_ in params: p_.grad = torch.zeros_like(p_)
opt = torch.optim.Adam(params, lr=1e-3, eps=1e-15, fused=True, capturable=True)
o, d = sc.rays(0); o, d = o[:1024].contiguous(), d[:1024].contiguous()
tgt = torch.rand((1024, 3), device=dev)
tup = mi.sampling_raytrace(d, o)
def probe(tag):
    with torch.no_grad():
        rgb, _ = rf(tup[0], d, ray_indices=tup[2])
    print(tag, float(rgb.mean()), float(rgb.std()), rf._handle_key)
probe("init")
gs = GraphedTrainStep(rf, opt, 1024, int(tup[0].shape[0]) + 64, mi.render_step_size)
gs.load(tup, d, tgt)
gs.capture(warmup=2)
probe("after capture")
for k in range(3):
    gs.load(tup, d, tgt); print("loss", float(gs.step()))
probe("after steps")
sc2 = scene.make_scene("smoke", device=dev)
with torch.no_grad():
    sc2.radiance_field.mlp_base.params.copy_(rf.mlp_base.params); sc2.radiance_field.mlp_head.params.copy_(rf.mlp_head.params)
    rgb2, _ = sc2.radiance_field(tup[0], d, ray_indices=tup[2])
print("fresh field with same params", float(rgb2.mean()), float(rgb2.std()))
