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


def band_rows(height: int, rank: int, world_size: int, band: int = 4) -> int:
    """Rows of the band-cyclic share of `rank`: the bands b (of `band` rows) with b % world_size == rank."""
    n_bands = (height + band - 1) // band
    mine = len(range(rank, n_bands, world_size))
    rows = mine * band
    if mine and (n_bands - 1) % world_size == rank:          # the last band may be short
        rows -= n_bands * band - height
    return rows


def assemble_banded(parts: Sequence[torch.Tensor], height: int, width: int, band: int = 4) -> torch.Tensor:
    """Inverse of the band-cyclic split: parts[r] is (band_rows(height, r, world) * width, C), rank r's bands in order;
    -> the (height * width, C) frame.  `height` must be a multiple of `band`."""
    world = len(parts)
    n_bands = height // band
    C = parts[0].shape[1]
    frame = parts[0].new_empty((n_bands, band * width, C))
    for r, p in enumerate(parts):
        frame[r::world] = p.view(-1, band * width, C)
    return frame.view(height * width, C)


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


class PeerFrame:
    """Frame buffers of a ray-sharded frame in the memory of ONE rank (`dst`), mapped into every other rank of the node over
    CUDA IPC (NVLink peer access): each rank's composite kernel stores its band-cyclic share of the pixels straight into
    them (`MeshRenderer.render_pose(bands=..., frame=...)`), so the final image gather is those stores — no gather
    collective, no concatenate / reassembly pass.  `sync()` is the only collective: a 4-byte all-reduce that orders every
    rank's stores before `dst` reads the frame (and `dst`'s reads of slot b before the peers' next stores to it).

    Collective constructor (every rank of the default process group calls it).  `slots` buffers rotate so that the ranks
    may run one frame ahead of the consumer.  Layout of a slot: rgb (H*W,3) | opacity (H*W,1) | depth (H*W,1), fp32."""

    def __init__(self, height: int, width: int, device, dst: int = 0, slots: int = 2):
        import ctypes as C
        from . import _lib
        self.H, self.W, self.dst, self.slots = height, width, dst, slots
        self.device = torch.device(device)
        self.rank, self.world = world()
        self._lib = lib = _lib.load()
        n = height * width
        self.slot_bytes = n * 5 * 4
        self._owned = self._mapped = None
        # Either every rank ends up with a usable mapping or every rank raises: a rank that fails never leaves its peers
        # waiting in a collective the others did not enter.
        base = C.c_void_p()
        err = None
        handle = None
        if self.rank == dst:
            try:
                _lib.check(lib.qf_peer_alloc(self.slot_bytes * slots, C.byref(base)), "qf_peer_alloc")
                self._owned = base.value
                if self.world > 1:
                    buf = C.create_string_buffer(64)
                    _lib.check(lib.qf_peer_export(C.c_void_p(self._owned), buf), "qf_peer_export")
                    handle = buf.raw
            except Exception as e:                       # noqa: BLE001 - reported to every rank below
                err = f"{type(e).__name__}: {e}"
        if self.world > 1:
            box = [(handle, err)]
            dist.broadcast_object_list(box, src=dst)
            handle, err = box[0]
            if err is None and self.rank != dst:
                try:
                    _lib.check(lib.qf_peer_open(handle, C.byref(base)), "qf_peer_open")
                    self._mapped = base.value
                except Exception as e:                   # noqa: BLE001
                    err = f"rank {self.rank}: {type(e).__name__}: {e}"
            ok = torch.tensor([0.0 if err else 1.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                self.close()
                raise RuntimeError(f"PeerFrame: peer mapping failed ({err or 'on another rank'})")
        elif err is not None:
            raise RuntimeError(f"PeerFrame: {err}")
        self.base = base.value
        self._flag = torch.zeros((1,), dtype=torch.float32, device=self.device)

    def pointers(self, slot: int):
        """(rgb, opacity, depth) device addresses of `slot` as seen from THIS rank."""
        n = self.H * self.W
        b = self.base + (slot % self.slots) * self.slot_bytes
        return b, b + n * 12, b + n * 16

    def frame(self, slot: int):
        """On `dst`: the (rgb (N,3), opacity (N,1), depth (N,1)) tensors viewing `slot` (no copy); None elsewhere."""
        if self.rank != self.dst:
            return None
        n = self.H * self.W

        class _Mem:      # __cuda_array_interface__ holder: lets torch view memory this library allocated
            pass
        out = []
        for ptr, shape in zip(self.pointers(slot), ((n, 3), (n, 1), (n, 1))):
            m = _Mem()
            m.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}
            out.append(torch.as_tensor(m, device=self.device))
        return tuple(out)

    def sync(self):
        """Stream-ordered fence over all ranks (current stream): after it, `dst` sees every rank's stores of the frames
        rendered before it."""
        if self.world > 1:
            dist.all_reduce(self._flag)

    def close(self):
        if self._mapped:
            torch.cuda.synchronize(self.device)
            self._lib.qf_peer_close(self._mapped)
            self._mapped = None
        if self.world > 1:
            dist.barrier()          # every mapping is closed before the owner frees the allocation
        if self._owned:
            torch.cuda.synchronize(self.device)
            self._lib.qf_peer_free(self._owned)
            self._owned = None


def all_reduce_gradients(params: Sequence[torch.Tensor], n_local, n_global=None) -> None:
    """Training mode: sum the per-rank gradients of the (replicated) hash table and MLPs — large tensors in place,
    the small ones in one flat bucket — and rescale a per-rank MEAN loss to the global sample count: every rank's
    gradient is weighted by n_local / n_global, so the result is the gradient of the mean over all ranks' samples
    (pass the number of samples the rank's loss averaged over, not its ray count, when the two differ).

    `n_local` / `n_global` may be Python numbers or 0-d device tensors; with `n_global=None` the counts are all-reduced
    on the device and the scale stays a device scalar (no host synchronisation).

    Every rank ALWAYS enters the same collectives in the same order: a parameter without `.grad` on this rank (zero-hit
    batch, unused branch) contributes zeros — an early return here would leave the peers blocked in NCCL."""
    rank, ws = world()
    if ws == 1:
        return
    params = [p for p in params if p.requires_grad and p.numel()]
    if not params:
        return
    dev = params[0].device
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    if n_global is None:
        n_loc = n_local.detach().to(device=dev, dtype=torch.float32).reshape(1) if isinstance(n_local, torch.Tensor) \
            else torch.tensor([float(n_local)], dtype=torch.float32, device=dev)
        total = n_loc.clone()
        dist.all_reduce(total)
        scale = n_loc / total.clamp_min(1.0)                 # device scalar; a rank with no samples scales its zeros by 0
    else:
        scale = float(n_local) / float(n_global)
    unit = isinstance(scale, float) and scale == 1.0
    # equal shares (scale == 1 / world) on NCCL: the collective averages by itself (ReduceOp.AVG) — no separate scaling pass
    # over the 50 MB table gradient (a 100 MB read-modify-write per step); gloo has no AVG and keeps the explicit scale
    avg = (isinstance(scale, float) and not unit and abs(scale * ws - 1.0) < 1e-12 and dist.get_backend() == "nccl")
    op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
    if avg:
        unit = True
    big = [g for g in grads if g.numel() >= (1 << 20) and g.is_contiguous()]
    small = [g for g in grads if not (g.numel() >= (1 << 20) and g.is_contiguous())]
    for g in big:                       # the hash table: reduced in place, no flatten / copy-back of tens of MB
        if not unit:
            g.mul_(scale)
        dist.all_reduce(g, op=op)
    if small:                           # MLP matrices and biases: one bucket, one launch-latency-bound collective
        flat = torch.cat([g.reshape(-1) for g in small])
        if not unit:
            flat.mul_(scale)
        dist.all_reduce(flat, op=op)
        torch._foreach_copy_(small, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in small]), small)])


def max_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
