"""A few quadrature-field training steps (for ncu captures of field_net_forward/backward_kernel).  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
from quadraturefields_b200.utils import train_field_step
from quadraturefields_b200.field import Field
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
n = 1 << 18
g = torch.Generator(device=dev).manual_seed(1)
o, d = sc.rays(0)
pi = torch.randint(0, sc.n_rays, (n,), device=dev, generator=g)
o, d = o[pi].contiguous(), d[pi].contiguous()
net = Field(scale=0.5, precision=16, log2_T=19, L=16, max_res=512, min_res=16, output_dim=1, hidden_size=16, num_features=2, back_prop=False, nl="elu").to(dev)
opt = torch.optim.Adam(list(net.parameters()), lr=2e-2, eps=1e-15)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    loss, m = train_field_step(net, sc.radiance_field, sc.mesh_intersect, o, d, opt)
torch.cuda.synchronize()
print("loss", float(loss), "samples", m)
