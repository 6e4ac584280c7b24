"""Diagnostic: work counters of the wide packet traversal (needs a -DQF_TRACE_STATS build of the library).
Run on the GPU box:  QF_EXTRA_NVCC_FLAGS=-DQF_TRACE_STATS python tools/diag_trace_stats.py c2 c2_dense c4"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import _lib, scene as S
lib = _lib.load()
dev = torch.device("cuda:0")
names = sys.argv[1:] or ["c2"]
for name in names:
    kw = {}
    base = name
    if name == "c2_dense":
        base, kw = "c2", dict(cam_radius=2.2)
    sc = S.make_scene(base, device=dev, build_field=(base != "c5"), **kw)
    o, d = sc.rays(1)
    st = (C.c_ulonglong * 8)()
    lib.qf_debug_trace_stats(st)
    if base == "c5":
        sc.render_baked(o, d, image_width=sc.W)
    else:
        sc.render(o, d, image_width=sc.W)
    torch.cuda.synchronize()
    lib.qf_debug_trace_stats(st)
    p = max(st[0], 1)
    live = max(st[7], 1)
    print(f"{name}: packets {st[0]} (beyond the root: {st[7]}), per live packet: node visits {(st[1]-st[0]+st[7])/live:.1f}, "
          f"pushes {st[6]/live:.1f}, triangle children iterated {st[2]/live:.1f}, with >=1 lane passing its box {st[3]/live:.1f}, "
          f"lanes passing per such triangle {st[4]/max(st[3],1):.1f}, hits per packet {st[5]/live:.1f}", flush=True)
    del sc
    torch.cuda.empty_cache()
