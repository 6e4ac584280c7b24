"""A few steps of bench.py's quadrature-field training legs (field_train: mesh path; field_train_occgrid: configs[2] with the
occupancy-grid marcher) on one GPU, for ncu captures of render_weights / field_net_* / occgrid_march / accumulate kernels."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
import bench
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
sc = S.make_scene("c2", device=dev)
args = argparse.Namespace(train_rays=1 << 18, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
barrier = lambda: torch.cuda.synchronize()
print({k: v for k, v in bench.run_field_train_steps(args, sc, dev, 0, 1, barrier).items() if k in ("ms_per_step", "value")})
print({k: v for k, v in bench.run_field_train_occgrid_steps(args, sc, dev, 0, 1, barrier).items() if k in ("ms_per_step", "value", "error")})
