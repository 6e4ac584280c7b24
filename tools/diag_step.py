"""Diagnostic: where does a bench step's time go (host overhead vs kernels)?  Run on the GPU box."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import _lib, scene as S
lib = _lib.load()
dev = torch.device("cuda:0")
sc = S.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c2", device=dev)
N = sc.n_rays
rays = [sc.rays(v) for v in range(16)]
out = dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev))
def run(tag, steps=20, prof=False, acc=False):
    lib.qf_profile_enable(1 if prof else 0)
    h = torch.zeros((), dtype=torch.int64, device=dev)
    for i in range(3):
        sc.render(*rays[i % 16], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(steps):
        sc.render(*rays[i % 16], out=out)
        if acc:
            h += out["n_hits"][0]
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    print(f"{tag}: event {e0.elapsed_time(e1)/steps:.3f} ms/step, host-enqueue {t_host/steps*1e3:.3f} ms/step, wall {t_wall/steps*1e3:.3f} ms/step")
    if prof:
        import ctypes as C
        ms3 = (C.c_double * 3)(); n = C.c_int64()
        lib.qf_profile_read(ms3, C.byref(n))
        print("   stages ms:", [x / max(n.value, 1) for x in ms3], "chunks", n.value)
    lib.qf_profile_enable(0)
run("plain")
run("plain2")
run("with hits_acc", acc=True)
run("with profile", prof=True)
run("plain3", steps=100)
