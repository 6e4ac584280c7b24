"""Diagnostic: the full-forward kernel variants (QF_FIELD_TC=0/1/2) timed on one frame's hit samples.  Run on the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
for cfg in ("c2", "c4"):
    sc = S.make_scene(cfg, device=dev)
    o, d = sc.rays(3)
    tup = sc.mesh_intersect.sampling_raytrace(d, o)
    pts, idx_ray = tup[0], tup[2]
    def run():
        with torch.no_grad():
            return sc.radiance_field(pts, d, ray_indices=idx_ray)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print("variant %s %s: %d hits, forward %.3f ms" % (os.environ.get("QF_FIELD_TC", "0"), cfg, pts.shape[0], e0.elapsed_time(e1) / 20), flush=True)
    del sc
'''
# arguments: "<QF_FIELD_TC>[:<QF_SHADE_WARPS>]"
for v in sys.argv[1:] or ["0", "1", "2"]:
    tc, _, warps = v.partition(":")
    env = dict(os.environ, QF_FIELD_TC=tc)
    if warps:
        env["QF_SHADE_WARPS"] = warps
    try:
        r = subprocess.run([sys.executable, "-c", CODE], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout.replace("variant", "variant[%s]" % v), r.stderr[-1500:] if r.returncode else "")
    except subprocess.TimeoutExpired:
        print("variant[%s] TIMEOUT" % v)
