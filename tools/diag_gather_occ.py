"""Diagnostic: the standalone hash-grid gather (no MLP) on one c2 frame's hit samples in the fused path's slot-major
order, as a function of resident warps per SM and of software pipelining.  Run on the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, torch, ctypes as C
sys.path.insert(0, os.getcwd())
from quadraturefields_b200 import scene as S, _lib
dev = torch.device("cuda:0")
sc = S.make_scene(os.environ.get("QF_DIAG_CFG", "c2"), device=dev)
o, d = sc.rays(3)
tup = sc.mesh_intersect.sampling_raytrace(d, o)
pts, r = tup[0], tup[2].long()
M = pts.shape[0]
_, x01 = sc.radiance_field.normalize(pts)
W = sc.W
px, py = r % W, r // W
tile = (py // 4) * (W // 8) + px // 8
intile = (py % 4) * 8 + px % 8
first = torch.ones_like(r, dtype=torch.bool); first[1:] = r[1:] != r[:-1]
start = torch.cummax(torch.where(first, torch.arange(M, device=dev), torch.zeros_like(r)), 0).values
slot = torch.arange(M, device=dev) - start
xt = x01[torch.argsort((tile * 64 + slot) * 32 + intile)].contiguous()
lib = _lib.load()
out = torch.empty(M, device=dev)
h = sc.radiance_field._native()
def run():
    lib.qf_debug_encode_sum(h, C.c_void_p(xt.data_ptr()), C.c_int64(M), C.c_void_p(out.data_ptr()), _lib.stream(dev))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("blocks/SM=%s threads=%s smem=%s pipe=%s: %d hits, gather %.3f ms" % (os.environ.get("QF_DEBUG_ENC_BLOCKS_PER_SM"), os.environ.get("QF_DEBUG_ENC_THREADS"), os.environ.get("QF_DEBUG_ENC_SMEM"), os.environ.get("QF_DEBUG_ENC_PIPE", "0"), M, e0.elapsed_time(e1) / 20), flush=True)
'''
# arguments: "<blocks per SM>[:<threads per block>[:<dynamic smem bytes>]]"
for pipe in ("0",):
    for arg in sys.argv[1:] or ["4", "5", "6", "8", "10", "12", "16"]:
        bps, _, rest = arg.partition(":")
        thr, _, smem = rest.partition(":")
        env = dict(os.environ, QF_DEBUG_ENC_BLOCKS_PER_SM=bps, QF_DEBUG_ENC_PIPE=pipe, QF_DEBUG_ENC_THREADS=thr or "128", QF_DEBUG_ENC_SMEM=smem or "0")
        r = subprocess.run([sys.executable, "-c", CODE], cwd=ROOT, env=env, capture_output=True, text=True, timeout=200)
        print(r.stdout.strip(), r.stderr[-800:] if r.returncode else "")
