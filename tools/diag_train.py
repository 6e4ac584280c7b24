"""A few training steps of the bench's train mode, for the ncu launch list."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import bench
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
args = argparse.Namespace(train_rays=1 << 18, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
def barrier(): torch.cuda.synchronize()
print(bench.run_train_steps(args, sc, dev, 0, 1, barrier))
