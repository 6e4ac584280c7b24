"""Does rendering alternate frames on two streams overlap the ALU-bound trace with the L1TEX-bound shading?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from quadraturefields_b200 import scene as S
dev = torch.device("cuda:0")
sc = S.make_scene("c2", device=dev)
N = sc.n_rays
rays = [sc.rays(v) for v in range(16)]
def run(n_streams, steps=40):
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    outs = [dict(rgb=torch.empty((N, 3), device=dev), opacity=torch.empty((N, 1), device=dev), depth=torch.empty((N, 1), device=dev)) for _ in range(n_streams)]
    hits = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(n_streams)]
    def go(k):
        for i in range(k):
            s = i % n_streams
            with torch.cuda.stream(streams[s]):
                sc.render(*rays[i % 16], out=outs[s], hits_out=hits[s], image_width=sc.W)
        for st in streams: torch.cuda.current_stream(dev).wait_stream(st)
    go(6); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for st in streams: st.wait_stream(torch.cuda.current_stream(dev))
    e0.record(); 
    for st in streams: st.wait_event(e0)
    go(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{n_streams} stream(s): {ms:.3f} ms/frame = {N/ms/1e6:.2f} Grays/s... ({N/(ms*1e-3)/1e9:.3f} Grays/s)", flush=True)
for n in (1, 2, 3, 1, 2):
    run(n)
