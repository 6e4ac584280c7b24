"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL; gloo on CPU for tests).

Rays are independent units: mesh/BVH, hash table, MLPs and textures are replicated on every
rank, and either whole frames (views) or contiguous ray bands of one frame are dealt to the ranks.  The only
inference collective is the final image gather; training adds an all-reduce of the flat parameter gradients.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin this process to the CPUs NVML reports as local to the GPU, so the pinned host buffers of the end-to-end path are
    first-touched on the GPU's NUMA node (8 ranks streaming rays in and images out otherwise cross the socket link).
    Returns the CPU list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may renumber devices: go through the UUID
            uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def view_for_step(step: int, rank: int, world_size: int, n_views: int) -> int:
    """Frame-sharded rendering: at every step the ranks render `world_size` consecutive views."""
    return (step * world_size + rank) % n_views


def shard_rays(n_rays: int, rank: int, world_size: int, image_width: int = 0) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a frame's rays for `rank`.  With `image_width` the cut points fall on 4-row
    bands so every shard keeps whole 8x4 tiles (the trace kernel's packet shape)."""
    unit = image_width * 4 if image_width and n_rays % (image_width * 4) == 0 else 1
    units = n_rays // unit
    base, rem = divmod(units, world_size)
    lo = (rank * base + min(rank, rem)) * unit
    hi = lo + (base + (1 if rank < rem else 0)) * unit
    return lo, hi


def gather_frame(parts: torch.Tensor, sizes: Sequence[int], dst: int = 0) -> Optional[torch.Tensor]:
    """Final image gather: rank r contributes `parts` (sizes[r], C); `dst` gets the (sum(sizes), C) frame."""
    rank, ws = world()
    if ws == 1:
        return parts
    # dist.gather needs equal shapes: pad every shard to the largest one, trim on the destination
    m = max(sizes)
    send = parts.contiguous()
    if send.shape[0] < m:
        send = torch.cat([send, send.new_zeros((m - send.shape[0],) + tuple(send.shape[1:]))], dim=0)
    out = [torch.empty_like(send) for _ in sizes] if rank == dst else None
    dist.gather(send, out, dst=dst)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0) if rank == dst else None


def all_reduce_gradients(params: Sequence[torch.Tensor], n_local: int, n_global: Optional[int] = None) -> None:
    """Training mode: sum the per-rank gradients of the (replicated) hash table and MLPs — large tensors in place,
    the small ones in one flat bucket — and rescale a per-rank mean loss to the global sample count (`n_local`
    samples here, `n_global` overall)."""
    rank, ws = world()
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    if ws == 1:
        return
    if n_global is None:
        counts = torch.tensor([float(n_local)], device=grads[0].device)
        dist.all_reduce(counts)
        n_global = float(counts.item())
    scale = float(n_local) / float(n_global)
    big = [g for g in grads if g.numel() >= (1 << 20) and g.is_contiguous()]
    small = [g for g in grads if not (g.numel() >= (1 << 20) and g.is_contiguous())]
    for g in big:                       # the hash table: reduced in place, no flatten / copy-back of tens of MB
        if scale != 1.0:
            g.mul_(scale)
        dist.all_reduce(g)
    if small:                           # MLP matrices and biases: one bucket, one launch-latency-bound collective
        flat = torch.cat([g.reshape(-1) for g in small])
        if scale != 1.0:
            flat.mul_(scale)
        dist.all_reduce(flat)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


def max_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
