"""Diagnostic: the field backward (qf_ngp_backward_inputs: absmax + ngp_backward_kernel + weight_grad_kernel + unpack) timed
alone on the hit samples of one 2^18-ray training batch of the c2 scene.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import _lib, scene as S
lib = _lib.load()
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
g = torch.Generator(device=dev).manual_seed(3)
n = 1 << 18
# random pixels of random views (train_finetune.py batches): incoherent rays
rays = [sc.rays(v) for v in range(8)]
O = torch.cat([r[0] for r in rays]); D = torch.cat([r[1] for r in rays])
pick = torch.randint(0, O.shape[0], (n,), generator=g, device=dev)
o, d = O[pick].contiguous(), D[pick].contiguous()
tup = sc.mesh_intersect.sampling_raytrace(d, o)
pts, idx_ray = tup[0].contiguous(), tup[2].contiguous()
M = pts.shape[0]
rf = sc.radiance_field
h = rf._native()
g_rgb = torch.randn((M, 3), device=dev, generator=g) * 1e-3
g_den = torch.randn((M,), device=dev, generator=g) * 1e-5
pb, ph = rf.mlp_base.params, rf.mlp_head.params
g_base, g_head = torch.zeros_like(pb), torch.zeros_like(ph)
nb = rf._n_base
ws = _lib.workspace(dev, lib.qf_ngp_backward_workspace_bytes(M), "ngp_bwd")
ri = idx_ray.to(torch.int64).contiguous()
def bwd():
    _lib.check(lib.qf_ngp_backward_inputs(h, _lib.ptr(pts), _lib.ptr(d), _lib.ptr(ri), M, _lib.ptr(g_rgb), _lib.ptr(g_den),
                                          _lib.ptr(g_base[nb:]), _lib.ptr(g_base[:nb]), _lib.ptr(g_head), None,
                                          _lib.ptr(ws), ws.numel(), _lib.stream(dev)), "qf_ngp_backward_inputs")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.no_grad():
    print(f"bwd_time: M={M} field backward {timeit(bwd):.3f} ms; forward {timeit(lambda: rf(pts, d, ray_indices=idx_ray)):.3f} ms", flush=True)
