"""Kernel-level breakdown of the training steps (torch.profiler / CUPTI): time per kernel and GPU idle share.  Run on the GPU box."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
from quadraturefields_b200.utils import render_train, train_field_step
from quadraturefields_b200.field import Field
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
n = 1 << 18
g = torch.Generator(device=dev).manual_seed(1)
pool = [sc.rays(v) for v in range(8)]
O_all, D_all = torch.stack([p[0] for p in pool]), torch.stack([p[1] for p in pool])
def batch():
    vi = torch.randint(0, 8, (n,), device=dev, generator=g); pi = torch.randint(0, sc.n_rays, (n,), device=dev, generator=g)
    return O_all[vi, pi].contiguous(), D_all[vi, pi].contiguous(), torch.rand((n, 3), device=dev, generator=g)
rf = sc.radiance_field
params = [rf.mlp_base.params, rf.mlp_head.params]
opt = torch.optim.Adam(params, lr=1e-4, eps=1e-15, fused=True)
net = Field(scale=0.5, precision=16, log2_T=19, L=16, max_res=512, min_res=16, output_dim=1, hidden_size=16, num_features=2, back_prop=False, nl="elu").to(dev)
fopt = torch.optim.Adam(list(net.parameters()), lr=2e-2, eps=1e-15)
for p_ in params: p_.grad = torch.zeros_like(p_)
for p_ in net.parameters(): p_.grad = torch.zeros_like(p_)
def step_rf(b):
    o, d, target = b
    opt.zero_grad(set_to_none=False)
    rgb, _, _, _ = render_train(sc.mesh_intersect, rf, o, d)
    torch.nn.functional.smooth_l1_loss(rgb, target).backward()
    opt.step()
def step_field(b):
    train_field_step(net, rf, sc.mesh_intersect, b[0], b[1], fopt)
for name, fn in (("radiance-field step", step_rf), ("quadrature-field step", step_field)):
    bs = [batch() for _ in range(6)]
    for b in bs[:3]: fn(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e0.record()
        for b in bs[3:]: fn(b)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    tot = sum(e.device_time_total for e in ev) / 3e3
    print(f"== {name}: {ms:.3f} ms/step, GPU busy {tot:.3f} ms/step ({100 * tot / ms:.0f} %)")
    for e in sorted(ev, key=lambda e: -e.device_time_total)[:14]:
        print(f"   {e.device_time_total / 3e3:7.3f} ms  x{e.count // 3:<3d} {e.key[:100]}")
