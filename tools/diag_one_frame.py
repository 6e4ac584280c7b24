"""Render a few frames of one config (for ncu captures of the large-config kernels).  usage: diag_one_frame.py <config> [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
baked = name.startswith("c5")
sc = S.make_scene(name, device=dev, build_field=not baked)
o, d = sc.rays(0)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    out = sc.render_baked(o, d, image_width=sc.W) if baked else sc.render(o, d, image_width=sc.W)
torch.cuda.synchronize()
print(name, int(out["n_hits"]))
