"""Diagnostic: host (enqueue) time against device time of the pose-in e2e loop of bench.py.  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
from quadraturefields_b200.utils import MeshRenderer
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
N, NB = sc.n_rays, int(os.environ.get("NB", "3"))
rs = [MeshRenderer(sc.mesh_intersect, radiance_field=sc.radiance_field) for _ in range(NB)]
d_out = [dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev)) for _ in range(NB)]
host = [dict(rgb=torch.empty((N, 3)).pin_memory(), depth=torch.empty((N, 1)).pin_memory()) for _ in range(NB)]
s_d2h = torch.cuda.Stream(dev); s_cmps = [torch.cuda.Stream(dev) for _ in range(NB)]
ev_cmp = [torch.cuda.Event() for _ in range(NB)]; ev_out = [torch.cuda.Event() for _ in range(NB)]
def loop(n, copy=True, render=True):
    for j in range(n):
        b = j % NB
        with torch.cuda.stream(s_cmps[b]):
            s_cmps[b].wait_event(ev_out[b])
            if render:
                rs[b].render_pose(sc.poses[j % len(sc.poses)], sc.W, sc.H, sc.focal, sc.cx, sc.cy, out=d_out[b])
            ev_cmp[b].record(s_cmps[b])
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_cmp[b])
            if copy:
                for k in ("rgb", "depth"):
                    host[b][k].copy_(d_out[b][k], non_blocking=True)
            ev_out[b].record(s_d2h)
for name, kw in (("render+copy", {}), ("render only", dict(copy=False)), ("copy only", dict(render=False))):
    loop(20, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(200, **kw); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: host enqueue {1e3*(t1-t0)/200:.3f} ms/frame, wall {1e3*(t2-t0)/200:.3f} ms/frame", flush=True)
