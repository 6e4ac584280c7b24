"""Run ONE of bench.py's ray-sharded legs on one GPU (profiling helper): python tools/run_leg.py c4|c5 [steps]."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
class A: steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
import __graft_entry__ as entry
entry.build()
res = bench.run_sharded_frame_leg(name, A, dev, 0, 1, lambda: torch.cuda.synchronize(dev), baked=(name == "c5"))
print(json.dumps(res))
