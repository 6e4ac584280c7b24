"""Stage times of the fused render on the larger BASELINE configs (c4 neural 1080p/1.15M tris/K=32, c5 baked 4K)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import _lib, scene as S
lib = _lib.load()
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["c4", "c5"]:
    baked = name.startswith("c5")
    sc = S.make_scene(name, device=dev, build_field=not baked)
    N = sc.n_rays
    rays = [sc.rays(v) for v in range(min(4, len(sc.poses)))]
    out = dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev))
    fn = (lambda o, d: sc.render_baked(o, d, out=out, image_width=sc.W)) if baked else (lambda o, d: sc.render(o, d, out=out, image_width=sc.W))
    for i in range(3): fn(*rays[i % len(rays)])
    torch.cuda.synchronize()
    lib.qf_profile_enable(1)
    ms3 = (C.c_double * 3)(); n = C.c_int64()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for i in range(steps): r = fn(*rays[i % len(rays)])
    e1.record(); torch.cuda.synchronize()
    lib.qf_profile_read(ms3, C.byref(n)); lib.qf_profile_enable(0)
    ms = e0.elapsed_time(e1) / steps
    hits = int(r["n_hits"])
    print(f"{name}: {N} rays, {sc.faces_np.shape[0]} tris, K={sc.K}, hits/ray {hits/N:.2f}: {ms:.3f} ms/frame = {N/(ms*1e-3)/1e6:.1f} Mrays/s; "
          f"stages/frame trace {ms3[0]/steps:.3f} shade {ms3[1]/steps:.3f} composite {ms3[2]/steps:.3f} ms; mesh {sc.mesh_intersect.rayintersector.info()}", flush=True)
    del sc, rays, out
    torch.cuda.empty_cache()
