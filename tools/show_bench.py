"""Print the key numbers of a bench.py JSON line (file argument or stdin)."""
import json, sys
text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads(text.strip().splitlines()[-1])
g = lambda k, kk="value": (d.get(k) or {}).get(kk)
print(f"n_gpus {d.get('n_gpus')} value {d['value']:.4g} ms/step {d['ms_per_step']:.4f} e2e {g('e2e'):.4g} png {g('e2e_png_ready') or 0:.4g} "
      f"rays-from-host {g('e2e_rays_from_host'):.4g} roofline {d['roofline']['frac']:.3f} stages {d.get('stage_ms_per_step')}")
for k in ("c2_dense", "c4", "c5"):
    v = d.get(k)
    if v:
        print(k, {kk: v.get(kk) for kk in ("ms_per_frame", "rank0_stage_ms_per_frame", "image_gather", "error") if v.get(kk) is not None})
for k in ("train", "field_train", "field_train_occgrid"):
    v = d.get(k)
    if v:
        print(k, v.get("ms_per_step"), v.get("value"))
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
