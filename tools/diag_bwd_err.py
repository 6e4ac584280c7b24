"""Diagnostic: per-matrix gradient error of qf_ngp_backward against autograd through the oracle (the setup of
tests/test_gpu_parity.py::test_ngp_backward_matches_oracle_autograd), for point subsets.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import quadfield_oracle as O
from quadraturefields_b200 import scene
from tests.helpers import oracle_params
T = torch.from_numpy
dev = torch.device("cuda:0")
sc = scene.make_scene("smoke", device=dev)
o, d = O.generate_rays(sc.poses[0], sc.W, sc.H, np.float32(sc.focal), np.float32(sc.cx), np.float32(sc.cy))
tup = O.sampling_raytrace(d, o, sc.vertices_np, sc.faces_np, sc.K)
g = torch.Generator().manual_seed(11)
x = torch.cat([T(tup[0]), (torch.rand(500, 3, generator=g) * 2 - 1) * 1.6])
dirs = torch.cat([T(d)[T(tup[2])], torch.nn.functional.normalize(torch.randn(500, 3, generator=g), dim=-1)])
M = x.shape[0]
wr, ws = torch.randn(M, 3, generator=g), torch.randn(M, 1, generator=g) * 0.01

def rel(a, b):
    a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)), float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))

for name, idx in (("all", slice(None)), ("hits", slice(0, M - 500)), ("random", slice(M - 500, M))):
    xs, ds, wrs, wss = x[idx], dirs[idx], wr[idx], ws[idx]
    p = oracle_params(sc)
    gp = lambda t: t.clone().requires_grad_()
    p = O.NGPParams(p.aabb, p.meta, gp(p.table), [gp(w) for w in p.base_w], [gp(w) for w in p.head_w])
    O.ROUND_HIDDEN = True
    rgb_r, den_r = O.ngp_forward(xs, ds, p)
    O.ROUND_HIDDEN = False
    ((rgb_r * wrs).sum() + (den_r * wss).sum()).backward()
    rf = sc.radiance_field
    rf.zero_grad(set_to_none=True)
    rgb, den = rf(xs.to(dev), ds.to(dev))
    print(name, "fwd rgb err", float((rgb.cpu() - rgb_r).abs().max()), "den rel", float(((den.cpu() - den_r).abs() / den_r.clamp_min(1e-3)).max()))
    ((rgb * wrs.to(dev)).sum() + (den * wss.to(dev)).sum()).backward()
    hw = rf.mlp_head.params.grad
    off = 0
    for k, w in enumerate(p.head_w):
        n = w.numel()
        print(name, "head", k, rel(hw[off:off + n], w.grad))
        if k == 0:
            got, ref = hw[off:off + n].view(64, 32).cpu(), w.grad
            colerr = (got - ref).abs().max(0).values / ref.abs().max()
            print("   per-column err/max:", [f"{float(c):.1e}" for c in colerr])
            print("   per-column |ref| max/global:", [f"{float(c):.2f}" for c in ref.abs().max(0).values / ref.abs().max()])
        off += n
    nb = rf._n_base
    print(name, "base", rel(rf.mlp_base.params.grad[:nb], torch.cat([w.grad.flatten() for w in p.base_w])))
    print(name, "table", rel(rf.mlp_base.params.grad[nb:], p.table.grad.flatten()))
