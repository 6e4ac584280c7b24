"""Diagnostic: trace/shade stage times for traversal modes x ray orders.  Run on the GPU box."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import _lib, scene as S
lib = _lib.load()
dev = torch.device("cuda:0")
sc = S.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c2", device=dev)
N = sc.n_rays
rays = [sc.rays(v) for v in range(8)]
out = dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev))
def run(tag, width, steps=10):
    lib.qf_profile_enable(1)
    for i in range(3):
        sc.render(*rays[i % 8], out=out, image_width=width)
    torch.cuda.synchronize()
    ms3 = (C.c_double * 3)(); n = C.c_int64()
    lib.qf_profile_read(ms3, C.byref(n))
    for i in range(steps):
        sc.render(*rays[i % 8], out=out, image_width=width)
    torch.cuda.synchronize()
    lib.qf_profile_read(ms3, C.byref(n))
    print(f"{tag}: trace {ms3[0]/n.value:.3f} shade {ms3[1]/n.value:.3f} composite {ms3[2]/n.value:.3f} ms  (mode={os.environ.get('QF_TRACE_MODE','0')})", flush=True)
    lib.qf_profile_enable(0)
run("strip32x1", 0)
run("tile8x4", sc.W)
