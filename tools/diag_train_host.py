"""Host-side profile of the bench's training legs (cProfile): where the Python time of a step goes.  Run on the GPU box."""
import sys, os, argparse, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import bench
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
args = argparse.Namespace(train_rays=1 << 18, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
def barrier(): torch.cuda.synchronize()
for fn in (bench.run_train_steps,):
    print(fn.__name__, "plain:", {k: v for k, v in fn(args, sc, dev, 0, 1, barrier).items() if k in ("value", "ms_per_step")})
    pr = cProfile.Profile()
    pr.enable()
    r = fn(args, sc, dev, 0, 1, barrier)
    pr.disable()
    print(fn.__name__, "under cProfile:", r["ms_per_step"], "ms/step")
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:]))
