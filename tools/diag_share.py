"""Diagnostic: ONE rank's band-cyclic share of a ray-sharded c4 / c5 frame on one GPU (what each of `world` ranks renders),
with per-stage times.  python tools/diag_share.py c4 8"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import _lib, scene as S
lib = _lib.load()
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
worlds = [int(a) for a in sys.argv[2:]] or [8]
band_rows = int(os.environ.get("BAND_ROWS", "4"))
baked = name == "c5"
sc = S.make_scene(name, device=dev, build_field=not baked)
r = sc.baked_renderer if baked else sc.renderer
for world in worlds:
    for rank in (0, world // 2):
        def frame(i):
            r.render_pose(sc.poses[i % len(sc.poses)], sc.W, sc.H, sc.focal, sc.cx, sc.cy, bands=(rank, world), band_rows=band_rows)
        for i in range(3): frame(i)
        torch.cuda.synchronize()
        lib.qf_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n): frame(3 + i)
        e1.record(); torch.cuda.synchronize()
        ms3 = (C.c_double * 3)(); nch = C.c_int64()
        lib.qf_profile_read(ms3, C.byref(nch)); lib.qf_profile_enable(0)
        print(f"share {name} band_rows={band_rows} world={world} rank={rank}: {e0.elapsed_time(e1)/n:.3f} ms per frame; trace {ms3[0]/n:.3f} shade {ms3[1]/n:.3f} composite {ms3[2]/n:.3f}", flush=True)
