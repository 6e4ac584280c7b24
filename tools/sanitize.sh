#!/bin/bash
# compute-sanitizer over the small configuration: smoke() (fused frame: wide-BVH trace, tcgen05 field kernel,
# composite), the three field-kernel variants, the baked path, the Embree-semantics trace and one training step.
# usage: tools/sanitize.sh <tag>      -> gpurun_out/sanitize_<tag>_{memcheck,racecheck,synccheck}.log
TAG=${1:-r2}
mkdir -p gpurun_out
cat > /tmp/qf_sanitize_driver.py <<'PY'
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as entry
entry.smoke()
from quadraturefields_b200 import scene
dev = torch.device("cuda:0")
sc = scene.make_scene("c5_small", device=dev)
o, d = sc.rays(0)
out = sc.render_baked(o, d, image_width=sc.W)
out2 = sc.render(o, d, image_width=sc.W)
# incoherent rays -> refill kernel; K=25 Embree semantics -> HitBufSmem + restart filter
perm = torch.randperm(o.shape[0], device=dev)
sc.mesh_intersect.rayintersector.set_restart_eps(3e-3)
tri, t, cnt = sc.mesh_intersect.rayintersector.trace(o[perm], d[perm], 25)
tri, t, cnt = sc.mesh_intersect.rayintersector.trace(o, d, 25)
sc.mesh_intersect.rayintersector.set_restart_eps(0.0)
# one training step (backward kernels, weight-gradient GEMM, table scatter)
from quadraturefields_b200.utils import render_train
rf = sc.radiance_field
rgb, _, _, n = render_train(sc.mesh_intersect, rf, o[perm][:4096], d[perm][:4096])
rgb.square().mean().backward()
torch.cuda.synchronize()
print("SANITIZE_DRIVER_OK", int(out["n_hits"]), int(out2["n_hits"]), int(cnt.sum()), n)
PY
for TOOL in memcheck racecheck synccheck; do
  for V in 1 0 2; do
    [ "$TOOL" != memcheck ] && [ "$V" != 1 ] && continue
    QF_FIELD_TC=$V timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/qf_sanitize_driver.py > gpurun_out/sanitize_${TAG}_${TOOL}_tc$V.log 2>&1
    echo "$TOOL QF_FIELD_TC=$V rc=$? $(grep -c 'SANITIZE_DRIVER_OK' gpurun_out/sanitize_${TAG}_${TOOL}_tc$V.log) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${TAG}_${TOOL}_tc$V.log | tail -1)"
  done
done
