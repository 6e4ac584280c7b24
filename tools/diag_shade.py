"""Diagnostic: standalone encode vs fused field kernel on one frame's real hit samples.  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
sc = S.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c2", device=dev)
o, d = sc.rays(3)
tup = sc.mesh_intersect.sampling_raytrace(d, o)
pts, idx_ray = tup[0], tup[2]
M = pts.shape[0]
sel, x01 = sc.radiance_field.normalize(pts)
x01 = x01.contiguous()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("hits", M)
import ctypes as C
from quadraturefields_b200 import _lib
lib = _lib.load()
outsum = torch.empty(M, device=dev)
h = sc.radiance_field._native()
def enc_sum(xx):
    lib.qf_debug_encode_sum(h, C.c_void_p(xx.data_ptr()), C.c_int64(M), C.c_void_p(outsum.data_ptr()), _lib.stream(dev))
print("gather only, 4 B out per sample (ray-major hits): %.3f ms" % timeit(lambda: enc_sum(x01)))
print("encode only (ray-major hits): %.3f ms" % timeit(lambda: sc.radiance_field.encode(x01)))
print("fused forward (ray-major hits): %.3f ms" % timeit(lambda: sc.radiance_field(pts, d, ray_indices=idx_ray)))
perm = torch.randperm(M, device=dev)
xs = x01[perm].contiguous()
print("encode only (shuffled hits): %.3f ms" % timeit(lambda: sc.radiance_field.encode(xs)))
# sorted by a coarse Morton-ish key (spatial coherence upper bound)
key = ((x01[:, 2] * 64).long() * 4096 + (x01[:, 1] * 64).long() * 64 + (x01[:, 0] * 64).long())
xm = x01[torch.argsort(key)].contiguous()
print("encode only (grid-sorted hits): %.3f ms" % timeit(lambda: sc.radiance_field.encode(xm)))
print("density only (ray-major): %.3f ms" % timeit(lambda: sc.radiance_field.query_density(pts)))
# slot-major inside 8x4 pixel tiles: the 32 lanes of a warp hold the j-th hits of 32 neighbouring pixels
W = sc.W
r = idx_ray.long()
px, py = r % W, r // W
tile = (py // 4) * (W // 8) + px // 8
intile = (py % 4) * 8 + px % 8
first = torch.ones_like(r, dtype=torch.bool); first[1:] = r[1:] != r[:-1]
start = torch.cummax(torch.where(first, torch.arange(M, device=dev), torch.zeros_like(r)), 0).values
slot = torch.arange(M, device=dev) - start
key2 = (tile * 64 + slot) * 32 + intile
xt = x01[torch.argsort(key2)].contiguous()
print("encode only (slot-major in 8x4 pixel tiles): %.3f ms" % timeit(lambda: sc.radiance_field.encode(xt)))
print("gather only (slot-major in 8x4 pixel tiles): %.3f ms" % timeit(lambda: enc_sum(xt)))
print("gather only (grid-sorted): %.3f ms" % timeit(lambda: enc_sum(xm)))
pt = pts[torch.argsort(key2)].contiguous(); dt = d[r[torch.argsort(key2)]].contiguous()
print("fused forward (slot-major in 8x4 pixel tiles): %.3f ms" % timeit(lambda: sc.radiance_field(pt, dt)))
