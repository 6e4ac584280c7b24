import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry
entry.build()
from oracle import quadfield_oracle as O
from quadraturefields_b200 import scene as S
from tests.helpers import oracle_params
dev = torch.device("cuda:0")
sc = S.make_scene("smoke", device=dev)
T = lambda a: torch.from_numpy(np.asarray(a))
o, d = O.generate_rays(sc.poses[0], sc.W, sc.H, np.float32(sc.focal), np.float32(sc.cx), np.float32(sc.cy))
tup = O.sampling_raytrace(d, o, sc.vertices_np, sc.faces_np, sc.K)
g = torch.Generator().manual_seed(11)
x = torch.cat([T(tup[0]), (torch.rand(500, 3, generator=g) * 2 - 1) * 1.6])
dirs = torch.cat([T(d)[T(tup[2])], torch.nn.functional.normalize(torch.randn(500, 3, generator=g), dim=-1)])
M = x.shape[0]
wr, ws = torch.randn(M, 3, generator=g), torch.randn(M, 1, generator=g) * 0.01
p0 = oracle_params(sc)
gr = lambda t: t.clone().requires_grad_()
p = O.NGPParams(p0.aabb, p0.meta, gr(p0.table), [gr(w) for w in p0.base_w], [gr(w) for w in p0.head_w])
O.ROUND_HIDDEN = bool(int(os.environ.get('RH','0')))
rgb_r, den_r = O.ngp_forward(x, dirs, p)
((rgb_r * wr).sum() + (den_r * ws).sum()).backward()
rf = sc.radiance_field
rf.zero_grad(set_to_none=True)
rgb, den = rf(x.to(dev), dirs.to(dev))
((rgb * wr.to(dev)).sum() + (den * ws.to(dev)).sum()).backward()
gh = rf.mlp_head.params.grad.cpu()
W3, W4, W5 = gh[:2048].view(64, 32), gh[2048:2048 + 4096].view(64, 64), gh[6144:].view(16, 64)
def rep(name, a, b):
    print(f"{name}: max|ref| {b.abs().max():.4g} max err {(a-b).abs().max():.4g} rel {(a-b).abs().max()/b.abs().max():.3g}")
rep("W3 all", W3, p.head_w[0].grad); rep("W3 sh cols", W3[:, :16], p.head_w[0].grad[:, :16]); rep("W3 feat cols", W3[:, 16:31], p.head_w[0].grad[:, 16:31]); rep("W3 pad col", W3[:, 31], p.head_w[0].grad[:, 31])
rep("W4", W4, p.head_w[1].grad); rep("W5 rows0-2", W5[:3], p.head_w[2].grad[:3]); rep("W5 rest", W5[3:], p.head_w[2].grad[3:] + 1e-30)
gb = rf.mlp_base.params.grad.cpu()
rep("W1", gb[:2048].view(64, 32), p.base_w[0].grad); rep("W2", gb[2048:3072].view(16, 64), p.base_w[1].grad)
rep("table", gb[3072:], p.table.grad.flatten())
