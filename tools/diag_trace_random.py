"""Time qf_trace_firstk on an incoherent ray batch (random pixels of random views) and on one coherent frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
sc = S.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c2", device=dev, build_field=False)
n = 1 << 18
g = torch.Generator(device=dev).manual_seed(1)
pool = [sc.rays(v) for v in range(16)]
O_all, D_all = torch.stack([p[0] for p in pool]), torch.stack([p[1] for p in pool])
vi = torch.randint(0, 16, (n,), device=dev, generator=g); pi = torch.randint(0, sc.n_rays, (n,), device=dev, generator=g)
o, d = O_all[vi, pi].contiguous(), D_all[vi, pi].contiguous()
ri = sc.mesh_intersect.rayintersector
def timeit(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print(f"random {n} rays K={sc.K}: {timeit(lambda: ri.trace(o, d, sc.K)):.3f} ms")
print(f"coherent frame {sc.n_rays} rays: {timeit(lambda: ri.trace(pool[0][0], pool[0][1], sc.K)):.3f} ms")
print(f"tuple (trace+scan+pack) random: {timeit(lambda: ri.trace_tuple(o, d, sc.K)):.3f} ms")
