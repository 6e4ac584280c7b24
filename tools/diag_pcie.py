"""Diagnostic: pinned host <-> device copy rates for the e2e frame payload (15.36 MB in, 12.8 MB out).  Run on the GPU box."""
import torch
dev = torch.device("cuda:0")
N = 640000
h_in = [torch.empty((N, 3)).pin_memory() for _ in range(2)]
d_in = [torch.empty((N, 3), device=dev) for _ in range(2)]
d_out = [torch.empty((N, k), device=dev) for k in (3, 1, 1)]
h_out = [torch.empty((N, k)).pin_memory() for k in (3, 1, 1)]
def t(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def h2d():
    for h, d in zip(h_in, d_in): d.copy_(h, non_blocking=True)
def d2h():
    for h, d in zip(h_out, d_out): h.copy_(d, non_blocking=True)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def both():
    with torch.cuda.stream(s1): h2d()
    with torch.cuda.stream(s2): d2h()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 15.36 MB: {a:.3f} ms = {15.36 / a:.1f} GB/s; D2H 12.8 MB: {b:.3f} ms = {12.8 / b:.1f} GB/s; both directions at once: {c:.3f} ms")
