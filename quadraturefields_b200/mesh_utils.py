"""Drop-in for the hot-path classes of the reference's `examples/mesh_utils.py`.

`RayIntersector` keeps the surface of the OptiX wrapper (mesh_utils.py:75-109) and
`MeshIntersection` that of mesh_utils.py:180-412 (`sampling_raytrace_numpy`, `sampling_indexing`,
`find_deltas`), but intersection runs on the GPU through libquadfield's LBVH + first-K traversal
(csrc/bvh.cu) instead of Embree in a DataLoader worker, and the depth re-sort stays on the device
instead of the reference's GPU->CPU `np.lexsort`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .mesh_io import Mesh as _Mesh, load_mesh  # noqa: F401  (OBJ with uv, ASCII / binary PLY)


class _DoneEvent:
    """Stands in for a CUDA event whose work the host has already waited for."""

    @staticmethod
    def synchronize():
        return None


class RayIntersector:
    """mesh_utils.py:75-109.  `intersects_id` returns (triangle_indices, ray_indices, psi) as numpy arrays."""

    def __init__(self, mesh, max_hits: int = 10, device="cuda", restart_eps: float = 0.0):
        self.mesh = mesh
        self.max_hits = max_hits
        self.device = torch.device(device)
        self._handle = None
        self._create(torch.as_tensor(np.asarray(mesh.vertices), dtype=torch.float32),
                     torch.as_tensor(np.asarray(mesh.faces), dtype=torch.int32))
        self.set_restart_eps(restart_eps)

    def set_restart_eps(self, eps: float):
        """0: all hits in (t, id) order, first K (the OptiX intersector's set).  > 0: the shipped trimesh + Embree
        intersector's set — a hit closer than `eps` to the previously kept one is skipped (include/quadfield.h)."""
        self.restart_eps = float(eps)
        _lib.check(_lib.load().qf_mesh_set_restart_eps(self._handle, C.c_float(self.restart_eps)), "qf_mesh_set_restart_eps")

    def _create(self, vertices: torch.Tensor, faces: torch.Tensor):
        lib = _lib.load()
        self.d_vertices = vertices.to(self.device, torch.float32).contiguous()
        self.d_faces = faces.to(self.device, torch.int32).contiguous()
        h = C.c_void_p()
        _lib.check(lib.qf_mesh_create(_lib.ptr(self.d_vertices), self.d_vertices.shape[0], _lib.ptr(self.d_faces),
                                      self.d_faces.shape[0], _lib.stream(self.device), C.byref(h)), "qf_mesh_create")
        self._handle = h

    @property
    def handle(self):
        return self._handle

    def info(self):
        lib = _lib.load()
        info = (C.c_int64 * 4)()
        pad = C.c_float()
        _lib.check(lib.qf_mesh_info(self._handle, info, C.byref(pad)), "qf_mesh_info")
        return dict(n_faces=info[0], n_vertices=info[1], n_nodes=info[2], device_bytes=info[3], box_pad=pad.value)

    def update_intersector(self, vertices):
        """mesh_utils.py:83-84 (`Intersector.update_vertices`): (V,3) vertex positions, same topology."""
        lib = _lib.load()
        v = torch.as_tensor(np.asarray(vertices) if not isinstance(vertices, torch.Tensor) else vertices)
        v = v.reshape(-1, 3).to(self.device, torch.float32).contiguous()
        if v.shape[0] != self.d_vertices.shape[0]:
            raise ValueError("update_intersector expects the (V,3) vertex array of the same topology")
        self.d_vertices = v
        _lib.check(lib.qf_mesh_update_vertices(self._handle, _lib.ptr(v), _lib.stream(self.device)), "qf_mesh_update_vertices")

    # ---- device-resident API --------------------------------------------------------------------
    @torch.no_grad()
    def trace(self, origins: torch.Tensor, vectors: torch.Tensor, max_hits: Optional[int] = None, with_total=False,
              with_t: bool = True):
        """First-K hits per ray: tri (N,K) int32 (-1 padded), t (N,K) (None with `with_t=False`: the tuple path recomputes
        the depth from the plane hit and skips the N*K floats), count (N,) [, total (N,)]."""
        lib = _lib.load()
        K = int(max_hits or self.max_hits)
        o = _lib.f32(origins, self.device)
        d = _lib.f32(vectors, self.device)
        N = o.shape[0]
        tri = torch.empty((N, K), dtype=torch.int32, device=self.device)
        t = torch.empty((N, K), dtype=torch.float32, device=self.device) if with_t else None
        count = torch.empty((N,), dtype=torch.int32, device=self.device)
        total = torch.empty((N,), dtype=torch.int32, device=self.device) if with_total else None
        ws = _lib.workspace(self.device, lib.qf_trace_workspace_bytes(N), "trace")
        _lib.check(lib.qf_trace_firstk(self._handle, _lib.ptr(o), _lib.ptr(d), N, K, _lib.ptr(tri), _lib.ptr(t),
                                       _lib.ptr(count), _lib.ptr(total), _lib.ptr(ws), ws.numel(), _lib.stream(self.device)),
                   "qf_trace_firstk")
        return (tri, t, count, total) if with_total else (tri, t, count)

    @torch.no_grad()
    def trace_tuple_begin(self, origins: torch.Tensor, vectors: torch.Tensor, max_hits: Optional[int] = None):
        """First half of `trace_tuple`: launch the traversal and the per-ray offset scan on the current stream and start
        an asynchronous copy of the hit total into pinned host memory.  No host synchronisation; hand the result to
        `trace_tuple_end` (on the same stream) once the size is wanted."""
        lib = _lib.load()
        K = int(max_hits or self.max_hits)
        o = _lib.f32(origins, self.device)
        d = _lib.f32(vectors, self.device)
        N = o.shape[0]
        tri, _, count = self.trace(o, d, K, with_t=False)
        offsets = torch.empty((N + 1,), dtype=torch.int64, device=self.device)
        ws = _lib.workspace(self.device, lib.qf_scan_workspace_bytes(N), "scan")
        st = _lib.stream(self.device)
        _lib.check(lib.qf_hits_offsets(_lib.ptr(count), N, _lib.ptr(offsets), _lib.ptr(ws), ws.numel(), st), "qf_hits_offsets")
        total = torch.empty((1,), dtype=torch.int64, pin_memory=True)
        total.copy_(offsets[N:], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return o, d, N, K, tri, count, offsets, total, ev

    @torch.no_grad()
    def tuple_buffers(self, capacity: int):
        """The six tensors of a hit tuple with room for `capacity` hits: (points, vectors, origins (cap,3) f32, depth (cap,)
        f32, index_ray, index_tri (cap,) i64)."""
        f = lambda *s: torch.empty((capacity,) + s, dtype=torch.float32, device=self.device)
        g = lambda: torch.empty((capacity,), dtype=torch.int64, device=self.device)
        return f(3), f(3), f(3), f(), g(), g()

    @torch.no_grad()
    def trace_tuple_end(self, pending, alloc=None):
        """Second half: wait (host) for the hit total only, size the tuple and pack it on the current stream.
        `alloc(M)` may supply `tuple_buffers` of capacity >= M (a prefetcher's ring); the tuple then aliases them."""
        lib = _lib.load()
        o, d, N, K, tri, count, offsets, total, ev = pending
        ev.synchronize()
        M = int(total[0])
        st = _lib.stream(self.device)
        # M changes from batch to batch: allocating in steps of 32 Ki hits lets the caching allocator hand the previous
        # batch's blocks straight back instead of growing the pool (cudaMalloc stalls the host for milliseconds)
        bufs = alloc(M) if alloc is not None else self.tuple_buffers((M + 32767) // 32768 * 32768)
        points, vecs, org, depth, index_ray, index_tri = (b[:M] for b in bufs)
        if M:
            _lib.check(lib.qf_hits_pack(self._handle, _lib.ptr(o), _lib.ptr(d), N, K, _lib.ptr(tri), _lib.ptr(count),
                                        _lib.ptr(offsets), _lib.ptr(points), _lib.ptr(vecs), _lib.ptr(index_ray),
                                        _lib.ptr(depth), _lib.ptr(index_tri), _lib.ptr(org), st), "qf_hits_pack")
        index_ray.qf_ray_major = True      # ascending by construction: `sampling_indexing` then skips its (host-synchronising) order check
        return points, vecs, index_ray, depth, index_tri, org, offsets

    @torch.no_grad()
    def trace_tuple(self, origins: torch.Tensor, vectors: torch.Tensor, max_hits: Optional[int] = None):
        """The reference data tuple on the device, ray-major and depth-sorted:
        (points (M,3), vectors (M,3), index_ray (M,), depth (M,), index_tri (M,), origins (M,3)); M may be 0."""
        lib = _lib.load()
        K = int(max_hits or self.max_hits)
        o = _lib.f32(origins, self.device)
        d = _lib.f32(vectors, self.device)
        N = o.shape[0]
        tri, _, count = self.trace(o, d, K, with_t=False)
        offsets = torch.empty((N + 1,), dtype=torch.int64, device=self.device)
        ws = _lib.workspace(self.device, lib.qf_scan_workspace_bytes(N), "scan")
        st = _lib.stream(self.device)
        _lib.check(lib.qf_hits_offsets(_lib.ptr(count), N, _lib.ptr(offsets), _lib.ptr(ws), ws.numel(), st), "qf_hits_offsets")
        total = C.c_int64()
        _lib.check(lib.qf_hits_total(_lib.ptr(offsets), N, C.byref(total), st), "qf_hits_total")
        pinned = torch.tensor([total.value], dtype=torch.int64)
        return self.trace_tuple_end((o, d, N, K, tri, count, offsets, pinned, _DoneEvent))

    # ---- reference surface ------------------------------------------------------------------------
    @torch.no_grad()
    def intersects_id(self, origins, vectors, multiple_hits=True, return_locations=True, max_hits=10):
        points, _, index_ray, depth, index_tri, _, _ = self.trace_tuple(torch.as_tensor(np.asarray(origins)),
                                                                      torch.as_tensor(np.asarray(vectors)),
                                                                      max_hits if multiple_hits else 1)
        return index_tri.cpu().numpy(), index_ray.cpu().numpy(), points.cpu().numpy()

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().qf_mesh_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


class MeshFinetune:
    """mesh_utils.py:112-156: accumulates the weight-averaged per-triangle displacement of the quadrature points and
    applies it to the vertices.  The caches, the vertices and the scatters stay on the device (`vertices_t`); the
    `vertices` property gives the reference's numpy view.  After `update_faces()` pass `vertices_t` to
    `RayIntersector.update_intersector` to refit the BVH (train_finetune.py:709-718)."""

    def __init__(self, vertices, faces, scaling, device="cuda") -> None:
        self.device = torch.device(device)
        self.vertices_t = torch.as_tensor(np.asarray(vertices, dtype=np.float32)).to(self.device).contiguous().clone()
        f = faces if isinstance(faces, torch.Tensor) else torch.as_tensor(np.asarray(faces))
        self.faces = f.to(self.device).long()
        self._faces32 = self.faces.to(torch.int32).contiguous()
        n = self.faces.shape[0]
        self.cache_d = torch.zeros((n, 3), device=self.device)
        self.cache_w = torch.ones(n, device=self.device) * 1e-8
        self.scaling = scaling

    @property
    def vertices(self) -> np.ndarray:
        return self.vertices_t.cpu().numpy()

    @torch.no_grad()
    def update_d(self, d, w, index_tri):
        lib = _lib.load()
        d = d.detach().to(self.device, torch.float32).contiguous()
        w = w.detach().to(self.device, torch.float32).contiguous()
        idx = index_tri.to(self.device, torch.int64).contiguous()
        if d.shape[0] != w.shape[0] or d.shape[0] != idx.shape[0] or d.dim() != 2 or d.shape[1] != 3:
            raise ValueError("update_d expects d (M,3), w (M,), index_tri (M,)")
        _lib.check(lib.qf_triangle_accumulate(_lib.ptr(d), _lib.ptr(w), _lib.ptr(idx), d.shape[0], self.cache_w.shape[0],
                                              _lib.ptr(self.cache_d), _lib.ptr(self.cache_w), _lib.stream(self.device)),
                   "qf_triangle_accumulate")

    @torch.no_grad()
    def update_faces(self):
        lib = _lib.load()
        V = self.vertices_t.shape[0]
        ws = _lib.workspace(self.device, lib.qf_vertex_displace_workspace_bytes(V))
        _lib.check(lib.qf_vertex_displace(_lib.ptr(self.cache_d), _lib.ptr(self.cache_w), _lib.ptr(self._faces32),
                                          self._faces32.shape[0], V, float(self.scaling), _lib.ptr(self.vertices_t),
                                          _lib.ptr(ws), ws.numel() * ws.element_size(), _lib.stream(self.device)),
                   "qf_vertex_displace")

    @torch.no_grad()
    def reset_d(self):
        self.cache_d[:] = 0
        self.cache_w[:] = 1e-8

    def sum_of_seq(self, gamma, epochs):
        s = 0
        for i in range(1, epochs):
            s += gamma ** i
        return s


@torch.no_grad()
def triangle_weight_max(triangles_weights: torch.Tensor, weights: torch.Tensor, index_tri: torch.Tensor) -> torch.Tensor:
    """The prune pass's `scatter_max(weights[:, 0], index_tri, out=zeros)` + `torch.maximum`
    (prune_mesh_after_finetuning.py:354-357), in place on `triangles_weights` (F,)."""
    lib = _lib.load()
    dev = triangles_weights.device
    if triangles_weights.dtype != torch.float32 or not triangles_weights.is_contiguous():
        raise ValueError("triangles_weights must be a contiguous float32 tensor")
    w = weights.detach().to(dev, torch.float32)
    w = (w[:, 0] if w.dim() == 2 else w).contiguous()
    idx = index_tri.to(dev, torch.int64).contiguous()
    _lib.check(lib.qf_triangle_weight_max(_lib.ptr(w), 1, _lib.ptr(idx), w.shape[0], triangles_weights.shape[0],
                                          _lib.ptr(triangles_weights), _lib.stream(dev)), "qf_triangle_weight_max")
    return triangles_weights


def _tag_ray_major(t: torch.Tensor) -> torch.Tensor:
    t.qf_ray_major = True
    return t


class HitTuple(tuple):
    """The reference's 7-tuple (points, vectors, index_ray, depth, index_tri, 0, origins) plus `.offsets`."""
    offsets = None


class MeshIntersection:
    """mesh_utils.py:180-412.  `mesh_path` may also be a `(vertices, faces)` pair."""

    def __init__(self, mesh_path, simplify_mesh=False, scale=1.0, num_repeat=16, optix=False, voxel_size=512,
                 num_intersections=20, render_step_size=0.005, device="cuda", hit_semantics: Optional[str] = None):
        if simplify_mesh:
            # the reference forces the trimesh loader (mesh_utils.py:186), for which simplification is not
            # available; every shipped script passes simplify_mesh=False
            raise NotImplementedError("simplify_mesh=True is not supported (no caller uses it)")
        self.mesh = load_mesh(mesh_path) if isinstance(mesh_path, str) else _Mesh(mesh_path[0], mesh_path[1])
        self.num_repeat = num_repeat
        self.num_intersections = num_intersections
        self.render_step_size = render_step_size
        self.mesh.vertices = self.mesh.vertices * scale                                   # mesh_utils.py:212
        self.device = torch.device(device)
        self.vertices = torch.from_numpy(self.mesh.vertices.astype(np.float32)).to(self.device)   # :214
        # Which hits a ray keeps.  The reference has two intersectors that disagree on closely spaced sheets:
        # `optix=True` (mesh_utils.py:203-221) returns all hits; `optix=False`, the shipped default (:223, :350-354), runs
        # trimesh's Embree loop, which restarts eps = clip(1e-4 * 100 / mesh.scale, 1e-8, inf) beyond every hit and so
        # never reports two hits closer than eps.  hit_semantics "all" | "embree"; None reads QF_HIT_SEMANTICS, default
        # "all": the set with a closed-form definition the oracle can brute-force; the two agree wherever consecutive
        # sheets are further than eps apart (every BASELINE config), and "embree" costs a 32-slot hit buffer per ray.
        import os
        if hit_semantics is None:
            hit_semantics = os.environ.get("QF_HIT_SEMANTICS", "all")
        if hit_semantics not in ("all", "embree"):
            raise ValueError("hit_semantics must be 'all' or 'embree'")
        self.hit_semantics = hit_semantics
        self.restart_eps = float(np.clip(1e-4 * 100.0 / float(self.mesh.scale), 1e-8, np.inf)) if hit_semantics == "embree" else 0.0
        self.rayintersector = RayIntersector(self.mesh, max_hits=self.num_intersections, device=device,
                                             restart_eps=self.restart_eps)

    def find_deltas(self, boundary, depth):
        """mesh_utils.py:225-231: the constant quadrature step (quirk Q4)."""
        deltas = torch.full((depth.shape[0],), self.render_step_size, dtype=torch.float32, device=depth.device)
        deltas.qf_const = float(self.render_step_size)      # read by utils.derive_properties instead of a device->host copy
        return deltas

    @torch.no_grad()
    def sampling_raytrace(self, vectors: torch.Tensor, origins: torch.Tensor):
        """Device-resident `sampling_raytrace_numpy`: returns the 7-tuple as CUDA tensors, or None on zero hits."""
        points, vecs, index_ray, depth, index_tri, org, offsets = self.rayintersector.trace_tuple(origins, vectors,
                                                                                                self.num_intersections)
        if index_tri.shape[0] == 0:
            return None
        tup = HitTuple((points, vecs, index_ray, depth, index_tri, 0, org))
        tup.offsets = offsets          # (N+1,) int64 start of every ray's run: lets callers skip re-deriving the packs
        return tup

    @torch.no_grad()
    def sampling_raytrace_begin(self, vectors: torch.Tensor, origins: torch.Tensor):
        """`sampling_raytrace` in two halves for prefetching: this one launches the intersection without waiting for it."""
        return self.rayintersector.trace_tuple_begin(origins, vectors, self.num_intersections)

    @torch.no_grad()
    def sampling_raytrace_end(self, pending, alloc=None):
        """Second half of `sampling_raytrace_begin`: the 7-tuple (or None), packed on the current stream."""
        points, vecs, index_ray, depth, index_tri, org, offsets = self.rayintersector.trace_tuple_end(pending, alloc)
        if index_tri.shape[0] == 0:
            return None
        tup = HitTuple((points, vecs, index_ray, depth, index_tri, 0, org))
        tup.offsets = offsets
        return tup

    def sampling_raytrace_numpy(self, vectors, origins, random=0):
        """mesh_utils.py:343-387 (numpy in, numpy out, `None` when nothing is hit)."""
        res = self.sampling_raytrace(torch.as_tensor(np.asarray(vectors)), torch.as_tensor(np.asarray(origins)))
        if res is None:
            return None
        points, vecs, index_ray, depth, index_tri, _, org = res
        return (points.cpu().numpy(), vecs.cpu().numpy(), index_ray.cpu().numpy(), depth.cpu().numpy(),
                index_tri.cpu().numpy(), 0, org.cpu().numpy())

    def sampling_indexing(self, points, origins, vectors, index_ray, depth, index_tri, random=0):
        """mesh_utils.py:389-412: per-ray re-sort by depth, pack boundaries, constant deltas — all on the device.  The
        permutation is found without gradient; the gathers that apply it are differentiable like the reference's
        `points[new_indices]` (the finetune step back-propagates through the re-sorted points)."""
        dev = self.device
        ray_major = bool(getattr(index_ray, "qf_ray_major", False))
        index_ray = _lib.i64(index_ray.to(dev))
        depth_in = depth.to(dev, torch.float32)
        with torch.no_grad():
            perm, boundary = self._resort(index_ray, _lib.f32(depth_in.detach(), dev), ray_major)
        g = lambda t: t.to(dev)[perm]
        depth = depth_in[perm]
        deltas = self.find_deltas(boundary, depth)
        return g(points), deltas, boundary, g(vectors), _tag_ray_major(index_ray[perm]), depth, g(index_tri), g(origins)

    def _resort(self, index_ray, depth, ray_major=False):
        """`ray_major=True` (index_ray produced by this package's tracer, tagged `qf_ray_major`) promises ascending ray ids
        and skips the order test, which is the only device->host read of the re-sort."""
        lib = _lib.load()
        dev = self.device
        M = index_ray.shape[0]
        if not ray_major and M > 1 and bool((index_ray[1:] < index_ray[:-1]).any()):
            # arbitrary order: full lexsort((depth, index_ray)) with device sorts (stable)
            o1 = torch.sort(depth, stable=True).indices
            perm = o1[torch.sort(index_ray[o1], stable=True).indices]
            boundary = torch.ones(M, dtype=torch.bool, device=dev)
            boundary[1:] = index_ray[perm][1:] != index_ray[perm][:-1]
        else:
            perm = torch.empty((M,), dtype=torch.int64, device=dev)
            b8 = torch.empty((M,), dtype=torch.uint8, device=dev)
            _lib.check(lib.qf_hits_resort(_lib.ptr(index_ray), _lib.ptr(depth), M, _lib.ptr(perm), _lib.ptr(b8),
                                          _lib.stream(dev)), "qf_hits_resort")
            boundary = b8.bool()
        return perm, boundary

