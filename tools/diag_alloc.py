"""Which torch.empty calls of a training step are slow (allocator growth, pinned host blocks)?  Run on the GPU box."""
import sys, os, argparse, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import bench
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
args = argparse.Namespace(train_rays=1 << 18, steps=20)
def barrier(): torch.cuda.synchronize()
bench.run_train_steps(args, sc, dev, 0, 1, barrier)
_empty = torch.empty
log = []
def timed_empty(*a, **k):
    t = time.perf_counter(); r = _empty(*a, **k); dt = time.perf_counter() - t
    log.append((dt, a, k.get("pin_memory", False), torch.cuda.current_stream().cuda_stream))
    return r
torch.empty = timed_empty
s0 = torch.cuda.memory_stats()
r = bench.run_train_steps(args, sc, dev, 0, 1, barrier)
s1 = torch.cuda.memory_stats()
torch.empty = _empty
print("ms/step", r["ms_per_step"], "cudaMalloc calls during run:", s1["num_device_alloc"] - s0["num_device_alloc"], "frees:", s1["num_device_free"] - s0["num_device_free"],
      "reserved MB", s1["reserved_bytes.all.current"] / 1e6)
print("total empty time ms:", 1e3 * sum(l[0] for l in log), "calls", len(log))
for dt, a, pin, st in sorted(log, key=lambda l: -l[0])[:25]:
    print(f"{dt * 1e3:8.3f} ms  size={a}  pin={pin} stream={st}")
