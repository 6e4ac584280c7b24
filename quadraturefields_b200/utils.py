"""Drop-in for the mesh-path pieces of the reference's `examples/utils.py`:
`derive_properties` (utils.py:863-898) and the mesh-path render drivers
`render_image_finetune_with_occgrid` (:465-607, inference: scaling=0 / no deformation field) and
`render_image_bake_texture_images_with_occgrid` (:998-1095).

Two levels are offered:
  * the reference's tuple-in / tuple-out functions (same arguments and return tuples), and
  * `MeshRenderer`, the fused frame render (rays -> image in three kernels, everything resident).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from .datasets.utils import Rays, namedtuple_map


def compress_sigma(sigma):
    """utils.py:54-58."""
    alpha = (1 - torch.exp(-sigma * 0.005))
    alpha = torch.clip(alpha * 255, 0, 255)
    return alpha.to(torch.uint8)


def inverse_of_compressed_sigma(alpha):
    """utils.py:60-63 (no clip — quirk Q5)."""
    alpha = alpha.to(torch.float32) / 255.0
    return -torch.log(1 - alpha) / 0.005


class _DerivePropertiesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, density, depths, delta, offsets, N, bg, bk):
        lib = _lib.load()
        dev = color.device
        M = density.shape[0]
        rgb = torch.empty((N, 3), dtype=torch.float32, device=dev)
        out_alpha = torch.empty((N, 1), dtype=torch.float32, device=dev)
        Depth = torch.empty((N, 1), dtype=torch.float32, device=dev)
        weights = torch.empty((M, 1), dtype=torch.float32, device=dev)
        _lib.check(lib.qf_derive_properties(_lib.ptr(color), _lib.ptr(density), _lib.ptr(depths), delta, _lib.ptr(offsets), N, bg,
                                            _lib.ptr(bk), _lib.ptr(rgb), _lib.ptr(out_alpha), _lib.ptr(Depth), _lib.ptr(weights),
                                            _lib.stream(dev)), "qf_derive_properties")
        ctx.save_for_backward(color, density, depths, offsets, bk)
        ctx.meta = (delta, N, bg)
        ctx.set_materialize_grads(False)
        return rgb, out_alpha, Depth, weights

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth, g_w):
        lib = _lib.load()
        color, density, depths, offsets, bk = ctx.saved_tensors
        delta, N, bg = ctx.meta
        M = density.shape[0]
        dev = color.device
        g_color, g_density = torch.empty_like(color), torch.empty_like(density)
        c = lambda t: _lib.f32(t) if t is not None else None
        if g_rgb is None and g_alpha is None and g_depth is None:
            g_rgb = torch.zeros((N, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.qf_derive_properties_backward(_lib.ptr(color), _lib.ptr(density), _lib.ptr(depths), delta, _lib.ptr(offsets),
                                                     N, M, bg, _lib.ptr(bk), _lib.ptr(c(g_rgb)), _lib.ptr(c(g_alpha)),
                                                     _lib.ptr(c(g_depth)), _lib.ptr(g_color), _lib.ptr(g_density),
                                                     _lib.stream(color.device)), "qf_derive_properties_backward")
        if g_w is not None and M:
            # the per-sample weights w_i = T_i (1 - e^{-sigma_i delta}) are an ordinary autograd output in the reference
            # (kaolin exponential_integration, utils.py:869-879): their gradient w.r.t. the densities is the nerfacc-style
            # weights backward with t_ends - t_starts = delta
            packed = torch.stack([offsets[:-1], offsets[1:] - offsets[:-1]], dim=1).contiguous()
            ts, te = torch.zeros_like(density), torch.full_like(density, float(delta))
            gin = torch.empty_like(density)
            _lib.check(lib.qf_render_weights_backward(1, _lib.ptr(density), _lib.ptr(ts), _lib.ptr(te), _lib.ptr(packed), N, M, None,
                                                      _lib.ptr(c(g_w.reshape(-1))), None, _lib.ptr(gin), _lib.stream(dev)),
                       "qf_render_weights_backward")
            g_density = g_density + gin
        return g_color, g_density, None, None, None, None, None, None


VALIDATE_INPUTS = bool(int(os.environ.get("QF_VALIDATE", "0")))   # debug: re-enable the input checks that synchronise the host


def derive_properties(color, density, depths, deltas, boundary, index_ray, render_bkgd=None, bg_color="white", N=0,
                      return_ray_ids=True):
    """utils.py:863-898 -> (rgb (N,3), out_alpha (N,1), index_ray[boundary], Depth (N,1), weights (M,1)).

    One kernel instead of 2 kaolin pack scans + 3 pack reductions + 3 scatters; differentiable w.r.t. color and
    density (all four float outputs, `weights` included).  `deltas` is the constant quadrature step (quirk Q4): a float, or
    the tensor `MeshIntersection.find_deltas` returns.  Packs are the runs of equal `index_ray`, ascending (the layout
    `sampling_indexing` produces); `boundary` is implied by `index_ray` and only used for the third return value.

    No device->host synchronisation happens here when `deltas` is a float / tagged tensor and `return_ray_ids=False`
    (the third return value `index_ray[boundary]` has a data-dependent length, so producing it must wait for the device —
    as it does in the reference; the package's own drivers do not ask for it).  `QF_VALIDATE=1` (or
    `utils.VALIDATE_INPUTS = True`) re-enables the checks that the step is constant and the packs ascend."""
    lib = _lib.load()
    dev = color.device
    color = _lib.f32(color.reshape(-1, 3))
    density = _lib.f32(density.reshape(-1))
    depths = _lib.f32(depths.reshape(-1))
    M = density.shape[0]
    index_ray = _lib.i64(index_ray)
    with torch.no_grad():
        if isinstance(deltas, torch.Tensor):
            const = getattr(deltas, "qf_const", None)
            if VALIDATE_INPUTS and M and not bool((deltas == deltas.reshape(-1)[0]).all()):
                raise NotImplementedError("derive_properties: the quadrature step is constant in every reference caller "
                                          "(mesh_utils.py:225-231)")
            delta = float(const) if const is not None else (float(deltas.reshape(-1)[0]) if M else 0.0)
        else:
            delta = float(deltas)
        if VALIDATE_INPUTS and M > 1 and bool((index_ray[1:] < index_ray[:-1]).any()):
            raise NotImplementedError("derive_properties expects ray-major (ascending index_ray) packs")
        # packs are runs of equal, ascending index_ray: per-ray counts + scan on the device (qf_pack_info), no host round trip
        offsets = torch.empty((N + 1,), dtype=torch.int64, device=dev)
        if N:
            packed = torch.empty((N, 2), dtype=torch.int64, device=dev)
            ws = _lib.workspace(dev, lib.qf_pack_info_workspace_bytes(N), "pack")
            _lib.check(lib.qf_pack_info(_lib.ptr(index_ray), M, N, _lib.ptr(packed), _lib.ptr(ws), ws.numel(), _lib.stream(dev)),
                       "qf_pack_info")
            offsets[:N] = packed[:, 0]
        offsets[N:] = M
        ids = index_ray[boundary] if return_ray_ids else None
        bk = _lib.f32(render_bkgd, dev) if render_bkgd is not None else None
    rgb, out_alpha, Depth, weights = _DerivePropertiesFn.apply(color, density, depths, delta, offsets, N,
                                                               _lib.BG_MODES.get(bg_color, 2), bk)
    return rgb, out_alpha, ids, Depth, weights


class MeshRenderer:
    """Fused mesh-path frame render (csrc/render.cu): BVH first-K trace -> compact hit records -> field or baked
    shading -> composite.  Rays, mesh, tables and textures stay resident; no host sync inside."""

    def __init__(self, mesh_intersect, radiance_field=None, compressor=None, uv=None, max_hits: Optional[int] = None,
                 render_step_size: Optional[float] = None):
        self.mi = mesh_intersect
        self.field = radiance_field
        self.compressor = compressor
        self.device = mesh_intersect.device
        self.K = int(max_hits or mesh_intersect.num_intersections)
        self.delta = float(render_step_size if render_step_size is not None else mesh_intersect.render_step_size)
        # private copy: the library caches per-triangle records keyed by this buffer's address (include/quadfield.h)
        self.uv = _lib.f32(uv, self.device).clone() if uv is not None else None
        self._hits = torch.zeros((1,), dtype=torch.int32, device=self.device)

    @torch.no_grad()
    def render(self, origins: torch.Tensor, viewdirs: torch.Tensor, bg_color="white", render_bkgd=None, out=None,
               hits_out: Optional[torch.Tensor] = None, image_width: Optional[int] = None, frame_out: Optional[tuple] = None):
        """-> dict(rgb (N,3), opacity (N,1), depth (N,1), n_hits (device int32 tensor)).
        `hits_out`: optional 1-element int32 CUDA tensor (e.g. a slot of a per-step buffer) receiving the hit count.
        `image_width`: the rays are a row-major image of that width ((H,W,3) inputs imply it) — a pure speed hint.
        `frame_out`: raw device addresses (rgb, opacity, depth) of frame buffers to store the image into instead of `out` —
        e.g. a slot of a `parallel.PeerFrame` in another GPU's memory (the final image gather as the composite kernel's own
        stores); needs `image_width` and whole 4-row bands.  -> None."""
        if image_width is None:
            image_width = origins.shape[1] if origins.dim() == 3 else 0
        if frame_out is not None:
            return self._render_to_frame(_lib.f32(origins.reshape(-1, 3), self.device), _lib.f32(viewdirs.reshape(-1, 3), self.device),
                                         bg_color, render_bkgd, hits_out, int(image_width), (4, 1, 0), frame_out)
        lib = _lib.load()
        dev = self.device
        o = _lib.f32(origins.reshape(-1, 3), dev)
        d = _lib.f32(viewdirs.reshape(-1, 3), dev)
        N = o.shape[0]
        if out is None:
            out = dict(rgb=torch.empty((N, 3), dtype=torch.float32, device=dev),
                       opacity=torch.empty((N, 1), dtype=torch.float32, device=dev),
                       depth=torch.empty((N, 1), dtype=torch.float32, device=dev))
        nbytes = lib.qf_render_workspace_bytes(N, self.K)
        ws = _lib.workspace(dev, nbytes, "render")
        bk = _lib.f32(render_bkgd, dev) if render_bkgd is not None else None
        bg = _lib.BG_MODES.get(bg_color, 2)
        st = _lib.stream(dev)
        mesh = self.mi.rayintersector.handle
        hits = self._hits if hits_out is None else hits_out
        if self.compressor is not None:
            _lib.check(lib.qf_render_mesh_baked(mesh, self.compressor.native(), _lib.ptr(self.uv), _lib.ptr(o), _lib.ptr(d), N,
                                                int(image_width), self.K, self.delta, bg, _lib.ptr(bk), _lib.ptr(out["rgb"]),
                                                _lib.ptr(out["opacity"]), _lib.ptr(out["depth"]), _lib.ptr(hits),
                                                _lib.ptr(ws), ws.numel(), st), "qf_render_mesh_baked")
        else:
            _lib.check(lib.qf_render_mesh_ngp(mesh, self.field._native(), _lib.ptr(o), _lib.ptr(d), N, int(image_width), self.K, self.delta, bg,
                                              _lib.ptr(bk), _lib.ptr(out["rgb"]), _lib.ptr(out["opacity"]),
                                              _lib.ptr(out["depth"]), _lib.ptr(hits), _lib.ptr(ws), ws.numel(), st),
                       "qf_render_mesh_ngp")
        out["n_hits"] = hits
        return out


    def frame_to_uint8(self, out: dict, rgb8: Optional[torch.Tensor] = None, depth8: Optional[torch.Tensor] = None,
                       with_depth: bool = True):
        """The images the reference's eval loops save (train_finetune.py:639-646): `rgb8 = uint8(clamp(rgb,0,1)*255)` (N,3)
        and `depth8 = uint8(depth / depth.max() * 255)` (N,), made on the device from a `render` dict — a quarter of the
        bytes of the fp32 frame to bring to the host.  -> (rgb8, depth8 or None), CUDA uint8 tensors."""
        lib = _lib.load()
        dev = self.device
        rgb, depth = out["rgb"], out["depth"]
        n = rgb.shape[0]
        if rgb8 is None:
            rgb8 = torch.empty((n, 3), dtype=torch.uint8, device=dev)
        if with_depth and depth8 is None:
            depth8 = torch.empty((n,), dtype=torch.uint8, device=dev)
        if not with_depth:
            depth8 = None
        if getattr(self, "_u8_scratch", None) is None:
            self._u8_scratch = torch.zeros((1,), dtype=torch.int32, device=dev)
        _lib.check(lib.qf_frame_to_u8(_lib.ptr(rgb), _lib.ptr(depth), n, _lib.ptr(rgb8), _lib.ptr(depth8), _lib.ptr(self._u8_scratch),
                                      _lib.stream(dev)), "qf_frame_to_u8")
        return rgb8, depth8

    def _render_to_frame(self, o, d, bg_color, render_bkgd, hits_out, W, band, ptrs):
        """The fused render of a band-cyclic share with its pixels stored into whole-frame buffers (`ptrs`: raw device
        addresses of rgb / opacity / depth, local or peer-mapped)."""
        lib = _lib.load()
        dev = self.device
        N = o.shape[0]
        if N == 0:
            return None
        ws = _lib.workspace(dev, lib.qf_render_workspace_bytes(N, self.K), "render")
        bk = _lib.f32(render_bkgd, dev) if render_bkgd is not None else None
        bg = _lib.BG_MODES.get(bg_color, 2)
        hits = self._hits if hits_out is None else hits_out
        mesh = self.mi.rayintersector.handle
        p_rgb, p_a, p_d = (C.c_void_p(p) for p in ptrs)
        if self.compressor is not None:
            _lib.check(lib.qf_render_mesh_baked_to_frame(mesh, self.compressor.native(), _lib.ptr(self.uv), _lib.ptr(o), _lib.ptr(d), N,
                                                         self.K, self.delta, bg, _lib.ptr(bk), band[0], band[1], band[2], int(W),
                                                         p_rgb, p_a, p_d, _lib.ptr(hits), _lib.ptr(ws), ws.numel(), _lib.stream(dev)),
                       "qf_render_mesh_baked_to_frame")
        else:
            _lib.check(lib.qf_render_mesh_ngp_to_frame(mesh, self.field._native(), _lib.ptr(o), _lib.ptr(d), N, self.K, self.delta, bg,
                                                       _lib.ptr(bk), band[0], band[1], band[2], int(W), p_rgb, p_a, p_d,
                                                       _lib.ptr(hits), _lib.ptr(ws), ws.numel(), _lib.stream(dev)),
                       "qf_render_mesh_ngp_to_frame")
        return None

    @torch.no_grad()
    def render_pose(self, c2w, W: int, H: int, focal: float, cx: float, cy: float, bg_color="white", render_bkgd=None, out=None,
                    hits_out: Optional[torch.Tensor] = None, opengl: bool = True, rows: Optional[tuple] = None,
                    bands: Optional[tuple] = None, frame=None, frame_slot: int = 0, band_rows: int = 4):
        """The evaluation frame of the reference from its real input: a 3x4 camera-to-world pose on the HOST.  The reference's
        `SubjectLoader.fetch_data` builds the W*H rays on the device from the pose (nerf_synthetic.py:289-378) and the eval
        loop renders them (train_finetune.py:586-617); here: `qf_generate_rays` (the 48-byte pose travels as kernel
        arguments, no host->device copy of rays) into resident scratch, then the fused render.
        `rows=(r0, r1)` renders only image rows [r0, r1) — a rank's contiguous band of a ray-sharded frame.
        `bands=(rank, world)` renders the rank's band-CYCLIC share instead: the `band_rows`-row bands b with b % world == rank, compact
        and in order (`parallel.assemble_banded` puts the shares back together); hit counts vary a lot over an image, and
        dealing the bands round-robin balances the ranks.  -> the `render` dict (of the rendered rows only).
        `frame` (with `bands`): a `parallel.PeerFrame`; the share's pixels are stored where they sit in slot `frame_slot` of
        the WHOLE frame, which lives on the gathering rank (peer memory over NVLink) — the image gather of a ray-sharded
        frame without a collective (`frame.sync()` afterwards orders the stores before the reader).  -> None."""
        lib = _lib.load()
        dev = self.device
        m = np.ascontiguousarray(np.asarray(c2w.cpu() if isinstance(c2w, torch.Tensor) else c2w, dtype=np.float32)[:3, :4])
        n = W * H
        key = ("pose_rays", n)
        buf = getattr(self, "_pose_rays", None)
        if buf is None or buf[0] != key:
            buf = (key, torch.empty((n, 3), dtype=torch.float32, device=dev), torch.empty((n, 3), dtype=torch.float32, device=dev))
            self._pose_rays = buf
        _, o, d = buf
        if bands is not None:
            rank, world = bands
            if band_rows < 4 or band_rows % 4 or H % band_rows:
                raise ValueError(f"band_rows={band_rows}: needs a multiple of 4 that divides the image height {H}")
            n_rows = int(lib.qf_band_rows(H, band_rows, world, rank))
            if n_rows < 0:
                raise ValueError(f"bands={bands}: needs 0 <= rank < world")
            _lib.check(lib.qf_generate_rays_banded(m.ctypes.data_as(C.POINTER(C.c_float)), W, H, float(focal), float(cx), float(cy),
                                                   1 if opengl else 0, band_rows, world, rank, _lib.ptr(o), _lib.ptr(d),
                                                   _lib.stream(dev)), "qf_generate_rays_banded")
            o, d = o[:n_rows * W], d[:n_rows * W]
            if frame is not None:
                return self._render_to_frame(o, d, bg_color, render_bkgd, hits_out, W, (band_rows, world, rank),
                                             frame.pointers(frame_slot))
            return self.render(o, d, bg_color=bg_color, render_bkgd=render_bkgd, out=out, hits_out=hits_out, image_width=W)
        _lib.check(lib.qf_generate_rays(m.ctypes.data_as(C.POINTER(C.c_float)), W, H, float(focal), float(cx), float(cy),
                                        1 if opengl else 0, _lib.ptr(o), _lib.ptr(d), _lib.stream(dev)), "qf_generate_rays")
        if rows is not None:
            o, d = o[rows[0] * W:rows[1] * W], d[rows[0] * W:rows[1] * W]
        return self.render(o, d, bg_color=bg_color, render_bkgd=render_bkgd, out=out, hits_out=hits_out, image_width=W)


class FramePipeline:
    """Renders successive frames on alternating CUDA streams.  The BVH traversal kernel is issue/ALU bound and the
    hash-grid shading kernel is L1-gather bound, so the trace of frame i+1 overlaps the shading of frame i when they
    run on different streams (measured on B200: 0.50 -> 0.43 ms per 800x800 frame).  Each stream has its own scratch."""

    def __init__(self, renderer: "MeshRenderer", n_streams: int = 2):
        self.renderer = renderer
        self.device = renderer.device
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self._hits = [torch.zeros((1,), dtype=torch.int32, device=self.device) for _ in range(n_streams)]   # one counter per stream
        self._i = 0

    def begin(self):
        """Order the pipeline after everything already queued on the current stream."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)

    def submit(self, origins, viewdirs, out=None, after: Optional[torch.cuda.Event] = None, **kw):
        """Queue one frame; returns (out dict, stream it runs on)."""
        k = self._i % len(self.streams)
        st = self.streams[k]
        self._i += 1
        kw.setdefault("hits_out", self._hits[k])
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(st):
            if after is not None:
                st.wait_event(after)
            res = self.renderer.render(origins, viewdirs, out=out, **kw)
        if out is None and res is not None:
            # allocated on the pipeline stream but consumed by the caller's stream after join(): tell the allocator
            for t in res.values():
                if isinstance(t, torch.Tensor):
                    t.record_stream(cur)
        return res, st

    def join(self):
        """Make the current stream wait for every queued frame."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)


class HitTuplePrefetcher:
    """Intersection of the NEXT training batches while the current one trains.  The reference does the same thing with
    DataLoader workers: `SubjectLoader.fetch_data` calls `mesh_intersect.sampling_raytrace_numpy` on the CPU (Embree) and
    hands the tuple over as `data["data"]` (train_finetune.py:494-509), a few batches ahead (torch's `prefetch_factor`).
    Here the trace runs on a side CUDA stream in two halves (`sampling_raytrace_begin/_end`): `submit` launches the
    traversal of the new batch WITHOUT waiting for it, after it has sized and packed the tuple of the batch submitted
    before — whose traversal was launched a whole training step earlier, so the host-side wait for the hit count (the
    tuple's size) is over by then.  `get` hands the tuples back in submission order.  Submit two batches before the first
    step and one more after each step's work has been launched: neither the host nor the training stream then ever waits
    for a traversal.  With a single batch in flight `get` completes it on the spot (the one-deep behaviour).  The
    intersection does not depend on the networks' parameters, so the result is identical to tracing inside the step."""

    def __init__(self, mesh_intersect, ring: int = 0):
        """`ring` > 0: the tuples live in `ring` buffer sets owned by the prefetcher instead of fresh allocations, so a
        steady-state training loop makes no allocator calls for them (blocks that crossed streams are handed back late by
        the caching allocator, the pool keeps growing for a while and every `cudaMalloc` stalls the host for
        milliseconds).  A tuple is then valid until `ring - 1` further `get()` calls; 4 suits the two-deep schedule."""
        self.mesh_intersect = mesh_intersect
        self.stream = torch.cuda.Stream(device=mesh_intersect.device)
        self._inflight = None            # (pending trace, origins, viewdirs): launched, tuple not sized yet
        self._ready = []                 # [(tuple, event, origins, viewdirs, ringed)] in submission order
        self.ring = int(ring)
        self._sets = [None] * self.ring          # buffer sets, slot = tuple number % ring
        self._got = [None] * self.ring           # main-stream event recorded by get() number g, at g % ring
        self._n_completed = 0
        self._n_got = 0

    def _ring_alloc(self, n):
        """alloc callback for tuple number n, or None when its slot's previous tuple may still be in use."""
        R = self.ring
        if R < 2 or (n >= R and self._n_got < n - R + 2):      # get(n-R+1) not called yet: tuple n-R is not retired
            return None
        if n >= R:
            self.stream.wait_event(self._got[(n - R + 1) % R])  # everything that consumed tuple n-R has been submitted before it
        main = torch.cuda.current_stream(self.mesh_intersect.device)     # (alloc itself runs under the side stream)

        def alloc(M):
            cur = self._sets[n % R]
            if cur is None or cur[0].shape[0] < M:
                if cur is not None:
                    for t in cur:
                        t.record_stream(main)
                cap = (int(M * 1.25) + 32767) // 32768 * 32768
                cur = self._sets[n % R] = self.mesh_intersect.rayintersector.tuple_buffers(max(cap, 32768))
            return cur
        return alloc

    def _complete(self):
        if self._inflight is None:
            return
        pending, origins, viewdirs = self._inflight
        self._inflight = None
        n = self._n_completed
        self._n_completed += 1
        alloc = self._ring_alloc(n) if self.ring else None
        with torch.cuda.stream(self.stream):
            tup = self.mesh_intersect.sampling_raytrace_end(pending, alloc)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._ready.append((tup, ev, origins, viewdirs, alloc is not None))

    @torch.no_grad()
    def submit(self, origins, viewdirs, rays_ready: bool = False):
        """`rays_ready=True`: the ray tensors were completed earlier (a data-loader batch), so the trace need not wait
        for the work queued on the training stream — that wait would serialise it behind the whole current step."""
        self._complete()                                    # pack the previous batch first: it must not queue behind this trace
        main = torch.cuda.current_stream(self.mesh_intersect.device)
        if not rays_ready:
            self.stream.wait_stream(main)                   # the rays may still be in flight on the training stream
        with torch.cuda.stream(self.stream):
            pending = self.mesh_intersect.sampling_raytrace_begin(viewdirs, origins)
        self._inflight = (pending, origins, viewdirs)

    def get(self):
        """-> the 7-tuple (or None when nothing was hit) of the oldest submitted batch, safe to use on the current stream."""
        if not self._ready:
            self._complete()
        tup, ev, origins, viewdirs, ringed = self._ready.pop(0)
        main = torch.cuda.current_stream(self.mesh_intersect.device)
        if self.ring:
            g = torch.cuda.Event()
            g.record(main)                                  # the consumers of all earlier tuples were submitted before this point
            self._got[self._n_got % self.ring] = g
        self._n_got += 1
        main.wait_event(ev)
        if tup is not None:
            # tensors allocated on the side stream: tell the allocator they are now in use on this one, or their memory
            # could be handed to the next prefetch while this step's backward still reads it (ring buffers are never freed;
            # the ray offsets always come from the allocator)
            for t in ([] if ringed else list(tup)) + [getattr(tup, "offsets", None)]:
                if isinstance(t, torch.Tensor):
                    t.record_stream(main)
        return tup


_NO_TUPLE = object()


def render_train(mesh_intersect, radiance_field, origins, viewdirs, bg_color="white", render_bkgd=None, tup=_NO_TUPLE):
    """Differentiable mesh-path render (train_finetune.py:494-509 with scaling=0): intersection (no gradient) ->
    field at the hits (gradients to the hash table and MLPs) -> derive_properties.  -> (rgb (N,3), opacity (N,1), depth (N,1), n_hits)."""
    N = origins.shape[0]
    if tup is _NO_TUPLE:                                    # otherwise: a tuple prefetched with HitTuplePrefetcher
        with torch.no_grad():
            tup = mesh_intersect.sampling_raytrace(viewdirs, origins)
    dev = mesh_intersect.device
    if tup is None:
        # nothing hit: the quirk-Q2 fill, still connected to the parameters (zero gradient) so that `loss.backward()` and
        # the gradient all-reduce run on this rank exactly as on its peers
        fill = 0.0 if bg_color == "black" else 1.0
        tie = sum((p.reshape(-1)[:1].sum() * 0.0 for p in radiance_field.parameters() if p.requires_grad and p.numel()),
                  torch.zeros((), device=dev))
        z = torch.zeros((N, 1), device=dev) + tie
        return torch.full((N, 3), fill, device=dev) + tie, z, z.clone(), 0
    points, _, index_ray, depth, _, _, _ = tup
    rgbs, sigmas = radiance_field(points, _lib.f32(viewdirs, dev), ray_indices=index_ray)
    offsets = getattr(tup, "offsets", None)
    if offsets is not None:
        # the intersector already knows where every ray's run starts: no boundary scan, no host synchronisation
        bk = _lib.f32(render_bkgd, dev) if render_bkgd is not None else None
        rgb, opacity, depth_img, _ = _DerivePropertiesFn.apply(_lib.f32(rgbs), _lib.f32(sigmas.reshape(-1)), _lib.f32(depth),
                                                               float(mesh_intersect.render_step_size), offsets, N,
                                                               _lib.BG_MODES.get(bg_color, 2), bk)
        return rgb, opacity, depth_img, points.shape[0]
    boundary = torch.ones_like(index_ray, dtype=torch.bool)
    boundary[1:] = index_ray[1:] != index_ray[:-1]
    rgb, opacity, _, depth_img, _ = derive_properties(rgbs, sigmas.squeeze(-1), depth, mesh_intersect.render_step_size, boundary,
                                                      index_ray, render_bkgd=render_bkgd, bg_color=bg_color, N=N, return_ray_ids=False)
    return rgb, opacity, depth_img, points.shape[0]


class GraphedTrainStep:
    """The optimiser step of the mesh-path training loop (train_finetune.py:494-531 with scaling=0: field at the hits ->
    derive_properties -> smooth-L1 against the pixels -> backward into the hash table and the MLPs -> Adam) replayed from
    two CUDA graphs instead of ~60 Python-driven launches: [zero_grad, forward, loss, backward] and [optimizer.step], with
    the gradient all-reduce (if any) issued eagerly between them.

    The obstacle to capturing the step is the data-dependent hit count M.  Everything runs on buffers of a fixed capacity C
    instead: the hits of the batch fill [0, M), the tail [M, C) holds dummy hits (origin of the field's box, depth 0) that
    belong to `n_dummy` extra rays N, N+1, ... of at most 8 hits each, whose colours reach no loss term — their gradient
    is exactly zero, the table scatter skips zero gradients and the weight-gradient GEMM adds zeros — so every kernel sees
    static shapes and the result equals the eager step on the M real hits.  `load()` (eager: a handful of copies) moves a
    traced tuple into the static buffers; a batch with more than C hits raises `OverflowError` (run it eagerly).

    `optimizer` must be capturable (`torch.optim.Adam(..., fused=True, capturable=True)`) and the parameters must already
    own `.grad` tensors; `radiance_field.accumulate_grad_in_place = True` is set."""

    def __init__(self, radiance_field, optimizer, n_rays: int, capacity: int, render_step_size: float, bg_color="white",
                 all_reduce=None):
        dev = radiance_field.aabb.device
        self.rf, self.opt, self.N, self.C, self.dev = radiance_field, optimizer, int(n_rays), int(capacity), dev
        self.delta, self.bg, self.all_reduce = float(render_step_size), _lib.BG_MODES.get(bg_color, 2), all_reduce
        self.D = (self.C + 7) // 8 + 1                                  # dummy rays: enough for an all-padding batch
        N, C, D = self.N, self.C, self.D
        self.points = torch.zeros((C, 3), device=dev)
        self.depth = torch.zeros((C,), device=dev)
        self.index_ray = torch.full((C,), N, dtype=torch.int64, device=dev)
        self.offsets = torch.zeros((N + D + 1,), dtype=torch.int64, device=dev)
        self.viewdirs = torch.zeros((N + D, 3), device=dev)
        self.viewdirs[N:, 2] = 1.0
        self.target = torch.zeros((N, 3), device=dev)
        self.loss = torch.zeros((), device=dev)
        self._pad_ray = N + torch.arange(C, device=dev) // 8             # ray of padding slot j (relative to M): N + j // 8
        self._pad_off = 8 * torch.arange(D + 1, device=dev)
        self.g_fb = self.g_opt = None
        radiance_field.accumulate_grad_in_place = True

    @torch.no_grad()
    def load(self, tup, viewdirs, target):
        """Copy a hit tuple (`MeshIntersection.sampling_raytrace` / the prefetcher), the batch's view directions (N,3) and
        target colours (N,3) into the static buffers and pad the tail."""
        N, C = self.N, self.C
        M = 0 if tup is None else int(tup[0].shape[0])
        if M > C:
            raise OverflowError(f"{M} hits exceed the captured capacity {C}")
        if M:
            self.points[:M].copy_(tup[0]); self.index_ray[:M].copy_(tup[2]); self.depth[:M].copy_(tup[3])
            self.offsets[:N + 1].copy_(tup.offsets)
        else:
            self.offsets[:N + 1].zero_()
        if M < C:
            self.points[M:].zero_(); self.depth[M:].zero_()
            self.index_ray[M:].copy_(self._pad_ray[:C - M])
        self.offsets[N:].copy_(torch.clamp(self._pad_off + M, max=C))
        self.viewdirs[:N].copy_(viewdirs)
        self.target.copy_(target)
        self.n_hits = M

    def _forward_backward(self):
        N = self.N
        self.opt.zero_grad(set_to_none=False)
        # the previous replay's optimiser step changed the parameters without touching Python's version counters (and a
        # capturable fused Adam does not bump them even when run eagerly): refresh the fp16 working copies explicitly, so
        # that the refresh kernels are part of the captured forward
        self.rf.mark_parameters_changed()
        rgbs, sigmas = self.rf(self.points, self.viewdirs, ray_indices=self.index_ray)
        rgb, _, _, _ = _DerivePropertiesFn.apply(_lib.f32(rgbs), _lib.f32(sigmas.reshape(-1)), self.depth, self.delta, self.offsets,
                                                 N + self.D, self.bg, None)
        loss = torch.nn.functional.smooth_l1_loss(rgb[:N], self.target)
        loss.backward()
        self.loss.copy_(loss.detach())

    def capture(self, warmup: int = 3):
        """Warm the step up on a side stream (allocations, lazily created kernel attributes), then capture the two graphs.
        The static buffers must hold a valid batch (`load`).  Parameters and optimiser state are restored afterwards (in
        place: the graphs hold their addresses), so capturing does not advance the training."""
        dev = self.dev
        params = [p for g in self.opt.param_groups for p in g["params"]]
        p_snap = [p.detach().clone() for p in params]
        s_snap = {id(p): {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self.opt.state.get(p, {}).items()} for p in params}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._forward_backward()
                if self.all_reduce is not None:
                    self.all_reduce()
                self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.g_fb, self.g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self._forward_backward()
        with torch.cuda.graph(self.g_opt, pool=self.g_fb.pool()):
            self.opt.step()
        with torch.no_grad():
            for p, p0 in zip(params, p_snap):
                p.copy_(p0)
                for k, v in self.opt.state.get(p, {}).items():
                    if isinstance(v, torch.Tensor):
                        old = s_snap[id(p)].get(k)
                        v.copy_(old) if isinstance(old, torch.Tensor) else v.zero_()
        self.rf.mark_parameters_changed()

    def step(self):
        """Replay one optimiser step on the loaded batch.  -> the loss (a device scalar that the next step overwrites)."""
        self.g_fb.replay()
        if self.all_reduce is not None:
            self.all_reduce()
        self.g_opt.replay()
        self.rf.mark_parameters_changed()   # the graph moved the parameters behind Python's version counters: the next eager
        return self.loss                    # use of the field must refresh its fp16 working copies


def field_training_batch(mesh_intersect, radiance_field, origins, viewdirs, tup=_NO_TUPLE):
    """The no-grad half of a quadrature-field training step (train_field.py:313-344): sample positions along the rays,
    the frozen radiance field's near-to-far and far-to-near weights at them (`rendering_field`, utils.py:431-446) and the
    normalised positions.  The reference takes its samples from nerfacc's occupancy-grid marcher (out
    of scope); here the sampler is the quadrature mesh: one interval of width render_step_size centred on every hit.
    -> (positions - 0.5 (M,3), dirs (M,3), weights (M,), weights_rev (M,)) or None when nothing is hit."""
    from .field_rendering import rendering_field
    with torch.no_grad():
        if tup is _NO_TUPLE:
            tup = mesh_intersect.sampling_raytrace(viewdirs, origins)
        if tup is None:
            return None
        points, _, index_ray, depth, _, _, _ = tup
        dev = points.device
        vd = _lib.f32(viewdirs, dev)
        half = 0.5 * float(mesh_intersect.render_step_size)
        t_starts, t_ends = depth - half, depth + half

        def rgb_sigma_fn(ts, te, ray_indices):
            rgbs, sigmas = radiance_field(points, vd, ray_indices=ray_indices)
            return rgbs, sigmas.squeeze(-1)

        _, _, _, weights, weights_rev = rendering_field(t_starts, t_ends, index_ray, n_rays=origins.shape[0], rgb_sigma_fn=rgb_sigma_fn)
        _, positions = radiance_field.normalize(points)                    # train_field.py:344
        return positions - 0.5, vd[index_ray], weights, weights_rev


def train_field_step(field_net, radiance_field, mesh_intersect, origins, viewdirs, optimizer, all_reduce=None, tup=_NO_TUPLE):
    """One optimiser step of the quadrature field (train_field.py:345-368): Field forward + analytic field_grad ->
    compute_field_loss -> backward (double backward through the MLP, first-order grid backward) -> optional gradient
    all-reduce -> optimizer.step().  -> (loss tensor, number of samples)."""
    batch = field_training_batch(mesh_intersect, radiance_field, origins, viewdirs, tup=tup)
    if batch is None:
        # a zero-hit batch still takes part in the step: with several ranks the peers are inside the gradient all-reduce,
        # so this rank contributes zero gradients instead of returning early (which would dead-lock NCCL)
        optimizer.zero_grad(set_to_none=False)
        if all_reduce is not None:
            all_reduce()
            optimizer.step()
        return None, 0
    positions, dirs, weights, weights_rev = batch
    _, field_grad = field_net(positions)
    loss = field_net.compute_field_loss(weights, weights_rev=weights_rev, field_norm=field_grad, view_dirs=dirs)
    optimizer.zero_grad(set_to_none=False)
    loss.backward()
    if all_reduce is not None:
        all_reduce()
    optimizer.step()
    return loss.detach(), positions.shape[0]


def _occgrid_fns(radiance_field, origins, viewdirs):
    """sigma_fn / rgb_sigma_fn of the reference's volumetric drivers (utils.py:93-122, 380-411): sample position =
    origin + dir * (t_start + t_end) / 2."""
    def positions(t_starts, t_ends, ray_indices):
        return origins[ray_indices] + viewdirs[ray_indices] * (t_starts + t_ends)[:, None] / 2.0

    def sigma_fn(t_starts, t_ends, ray_indices):
        return radiance_field.query_density(positions(t_starts, t_ends, ray_indices)).squeeze(-1)

    def rgb_sigma_fn(t_starts, t_ends, ray_indices):
        rgbs, sigmas = radiance_field(positions(t_starts, t_ends, ray_indices), viewdirs, ray_indices=ray_indices)
        return rgbs, sigmas.squeeze(-1)
    return sigma_fn, rgb_sigma_fn


def render_image_with_occgrid(radiance_field, estimator, rays: Rays, near_plane: float = 0.0, far_plane: float = 1e10,
                              render_step_size: float = 1e-3, render_bkgd: Optional[torch.Tensor] = None, cone_angle: float = 0.0,
                              alpha_thre: float = 0.0, test_chunk_size: int = 8192, timestamps=None, use_eps_loss: bool = False):
    """utils.py:65-172: the volumetric (ray-marched) render of the NeRF stage and of the finetune step's second pass:
    `estimator.sampling` (occupancy-grid marcher) -> `rendering` (nerfacc-style compositing).  -> (rgb, opacity, depth,
    n_rendering_samples, extras)."""
    from .field_rendering import rendering
    if timestamps is not None:
        raise NotImplementedError("dnerf timestamps are not part of the Quadfield scripts")
    rays, rays_shape, num_rays = _flatten_rays(rays)
    dev = estimator.binaries.device
    chunk = torch.iinfo(torch.int32).max if radiance_field.training else test_chunk_size
    results, extras = [], {}
    for i in range(0, num_rays, chunk):
        origins, viewdirs = _lib.f32(rays.origins[i:i + chunk], dev), _lib.f32(rays.viewdirs[i:i + chunk], dev)
        sigma_fn, rgb_sigma_fn = _occgrid_fns(radiance_field, origins, viewdirs)
        ray_indices, t_starts, t_ends = estimator.sampling(origins, viewdirs, sigma_fn=sigma_fn, near_plane=near_plane,
                                                           far_plane=far_plane, render_step_size=render_step_size,
                                                           stratified=radiance_field.training, cone_angle=cone_angle,
                                                           alpha_thre=alpha_thre)
        rgb, opacity, depth, extras = rendering(t_starts, t_ends, ray_indices, n_rays=origins.shape[0], rgb_sigma_fn=rgb_sigma_fn,
                                                render_bkgd=render_bkgd)
        extras["t_starts"], extras["t_ends"], extras["ray_indices"], extras["t_origins"] = t_starts, t_ends, ray_indices, origins
        results.append([rgb, opacity, depth, len(t_starts)])
    colors, opacities, depths, n_rendering_samples = [torch.cat(r, dim=0) if isinstance(r[0], torch.Tensor) else r
                                                      for r in zip(*results)]
    return (colors.view((*rays_shape[:-1], -1)), opacities.view((*rays_shape[:-1], -1)), depths.view((*rays_shape[:-1], -1)),
            sum(n_rendering_samples), extras)


@torch.no_grad()
def render_image_with_occgrid_test(max_samples: int, radiance_field, estimator, rays: Rays, near_plane: float = 0.0,
                                   far_plane: float = 1e10, render_step_size: float = 1e-3,
                                   render_bkgd: Optional[torch.Tensor] = None, cone_angle: float = 0.0, alpha_thre: float = 0.0,
                                   early_stop_eps: float = 1e-4, timestamps=None):
    """utils.py:175-350 (imported by train_field.py:14): the evaluation render that marches every ray a few samples at a
    time — `n_samples = max(min(num_rays // n_alive, 64), min_samples)` per round, the marcher capped at that many samples
    per ray and resumed from its termination planes — composites each round on top of the accumulated opacity
    (`prefix_trans`), and retires rays whose opacity passed 1 - early_stop_eps or that left the grid.
    -> (rgb, opacity, depth (un-normalised, :343), total samples, positions of every evaluated sample)."""
    from .field_rendering import accumulate_along_rays_, render_weight_from_density
    if timestamps is not None:
        raise NotImplementedError("dnerf timestamps are not part of the Quadfield scripts")
    rays, rays_shape, num_rays = _flatten_rays(rays)
    dev = estimator.binaries.device
    origins, viewdirs = _lib.f32(rays.origins, dev), _lib.f32(rays.viewdirs, dev)
    opacity = torch.zeros((num_rays, 1), device=dev)
    depth = torch.zeros((num_rays, 1), device=dev)
    rgb = torch.zeros((num_rays, 3), device=dev)
    ray_mask = torch.ones((num_rays,), dtype=torch.bool, device=dev)
    min_samples = 1 if cone_angle == 0 else 4              # 1 for synthetic scenes, 4 for real scenes (:229)
    iter_samples = total_samples = 0
    near_planes = torch.full((num_rays,), float(near_plane), dtype=torch.float32, device=dev)
    opc_thre = 1 - early_stop_eps
    positions_all = []
    while iter_samples < max_samples:
        n_alive = int(ray_mask.sum().item())               # the reference synchronises here too (:254)
        if n_alive == 0:
            break
        n_samples = max(min(num_rays // n_alive, 64), min_samples)
        iter_samples += n_samples
        ray_indices, t_starts, t_ends, offsets, termination = estimator.march(
            origins, viewdirs, near_planes, near_plane, far_plane, render_step_size, cone_angle, max_samples=n_samples,
            ray_mask=ray_mask, return_termination=True)
        counts = offsets[1:] - offsets[:-1]
        positions = origins[ray_indices] + viewdirs[ray_indices] * (t_starts + t_ends)[:, None] / 2.0
        positions_all.append(positions)
        if positions.shape[0]:
            rgbs, sigmas = radiance_field(positions, viewdirs, ray_indices=ray_indices)
            weights, _, alphas = render_weight_from_density(t_starts, t_ends, sigmas.squeeze(-1), ray_indices=ray_indices,
                                                            n_rays=num_rays, prefix_trans=1 - opacity[ray_indices].squeeze(-1))
            ri, ts, te = ray_indices, t_starts, t_ends
            if alpha_thre > 0:
                vis = alphas >= alpha_thre
                ri, rgbs, weights, ts, te = ri[vis], rgbs[vis], weights[vis], ts[vis], te[vis]
            accumulate_along_rays_(weights, values=rgbs, ray_indices=ri, outputs=rgb)
            accumulate_along_rays_(weights, values=None, ray_indices=ri, outputs=opacity)
            accumulate_along_rays_(weights, values=(ts + te)[..., None] / 2.0, ray_indices=ri, outputs=depth)
            total_samples += int(ri.shape[0])
        near_planes = termination
        ray_mask = torch.logical_and(opacity.view(-1) <= opc_thre, counts == n_samples)
    bk = _lib.f32(render_bkgd, dev) if render_bkgd is not None else torch.zeros(3, device=dev)
    rgb = rgb + bk * (1.0 - opacity)
    pos = torch.cat(positions_all, dim=0) if positions_all else torch.zeros((0, 3), device=dev)
    return (rgb.view((*rays_shape[:-1], -1)), opacity.view((*rays_shape[:-1], -1)), depth.view((*rays_shape[:-1], -1)),
            total_samples, pos)


def render_image_field_with_occgrid(radiance_field, estimator, rays: Rays, near_plane: float = 0.0, far_plane: float = 1e10,
                                    render_step_size: float = 1e-3, render_bkgd: Optional[torch.Tensor] = None,
                                    cone_angle: float = 0.0, alpha_thre: float = 0.0, test_chunk_size: int = 8192, timestamps=None):
    """utils.py:353-462 (the no-grad half of train_field.py's step): marcher samples, `rendering_field` -> near-to-far and
    far-to-near weights, sample positions and directions.  -> (rgb, opacity, depth, n_samples, weights, weights_rev,
    positions, dirs)."""
    from .field_rendering import rendering_field
    if timestamps is not None:
        raise NotImplementedError("dnerf timestamps are not part of the Quadfield scripts")
    rays, rays_shape, num_rays = _flatten_rays(rays)
    dev = estimator.binaries.device
    chunk = torch.iinfo(torch.int32).max if radiance_field.training else test_chunk_size
    results = []
    for i in range(0, num_rays, chunk):
        origins, viewdirs = _lib.f32(rays.origins[i:i + chunk], dev), _lib.f32(rays.viewdirs[i:i + chunk], dev)
        sigma_fn, rgb_sigma_fn = _occgrid_fns(radiance_field, origins, viewdirs)
        ray_indices, t_starts, t_ends = estimator.sampling(origins, viewdirs, sigma_fn=sigma_fn, near_plane=near_plane,
                                                           far_plane=far_plane, render_step_size=render_step_size,
                                                           stratified=radiance_field.training, cone_angle=cone_angle,
                                                           alpha_thre=alpha_thre, early_stop_eps=1e-4)   # utils.py:432
        rgb, opacity, depth, weights, weights_rev = rendering_field(t_starts, t_ends, ray_indices, n_rays=origins.shape[0],
                                                                    rgb_sigma_fn=rgb_sigma_fn, render_bkgd=render_bkgd)
        positions = origins[ray_indices] + viewdirs[ray_indices] * (t_starts + t_ends)[:, None] / 2.0
        results.append([rgb, opacity, depth, len(t_starts), weights, weights_rev, positions, viewdirs[ray_indices]])
    colors, opacities, depths, n_rendering_samples, weights, weights_rev, positions, dirs = [
        torch.cat(r, dim=0) if isinstance(r[0], torch.Tensor) else r for r in zip(*results)]
    return (colors.view((*rays_shape[:-1], -1)), opacities.view((*rays_shape[:-1], -1)), depths.view((*rays_shape[:-1], -1)),
            sum(n_rendering_samples), weights, weights_rev, positions, dirs)


def train_field_step_occgrid(field_net, radiance_field, estimator, rays: Rays, optimizer, near_plane=0.0, render_step_size=5e-3,
                             cone_angle=0.0, alpha_thre=0.0, render_bkgd=None):
    """train_field.py:320-368 with the reference's own sampler: frozen radiance field + occupancy-grid marcher ->
    weights / reversed weights -> quadrature Field forward with field_grad -> field loss -> backward -> optimizer step."""
    with torch.no_grad():
        _, _, _, n, weights, weights_rev, positions, dirs = render_image_field_with_occgrid(
            radiance_field, estimator, rays, near_plane=near_plane, render_step_size=render_step_size, render_bkgd=render_bkgd,
            cone_angle=cone_angle, alpha_thre=alpha_thre)
        if n == 0:
            return None, 0
        _, positions = radiance_field.normalize(positions)                 # train_field.py:344
    positions = positions - 0.5
    _, field_grad = field_net(positions)
    loss = field_net.compute_field_loss(weights, weights_rev=weights_rev, field_norm=field_grad, view_dirs=dirs)
    optimizer.zero_grad(set_to_none=False)
    loss.backward()
    optimizer.step()
    return loss.detach(), n


def train_finetune_step(radiance_field, field_net, estimator, mesh_intersect, mesh_finetune, rays: Rays, data, pixels, optimizer,
                        scaling=1 / 128, near_plane=0.0, render_step_size=5e-3, render_bkgd=None, cone_angle=0.0, alpha_thre=0.0,
                        bg_color="white", all_reduce=None):
    """One step of the finetune loop (train_finetune.py:494-533): the mesh-path render with the deformation field
    (discrete loss + regulariser, MeshFinetune accumulation), the volumetric render through the occupancy grid (smooth
    loss), both smooth-L1 against the pixels, backward into the radiance field AND the deformation field, optimizer step.
    `data` is the hit tuple of the batch (xyzs, dirs, index_ray, ts, index_tri, origins).
    -> dict(loss, rgb_discrete_loss, rgb_smooth_loss, loss_reg, n_mesh_samples, n_volume_samples)."""
    F = torch.nn.functional
    rgb, _, _, n_mesh, _, _, _, loss_reg, _ = render_image_finetune_with_occgrid(
        radiance_field, field_net, estimator, rays, data, near_plane=near_plane, render_step_size=render_step_size,
        render_bkgd=render_bkgd, cone_angle=cone_angle, alpha_thre=alpha_thre, mesh_intersect=mesh_intersect,
        mesh_finetune=mesh_finetune, scaling=scaling, bg_color=bg_color)
    rgb_full, _, _, n_vol, _ = render_image_with_occgrid(radiance_field, estimator, rays, near_plane=near_plane,
                                                       render_step_size=render_step_size, render_bkgd=render_bkgd,
                                                       cone_angle=cone_angle, alpha_thre=alpha_thre)
    pixels = pixels.to(rgb.device)
    rgb_discrete_loss = F.smooth_l1_loss(rgb.reshape(-1, 3), pixels.reshape(-1, 3))
    rgb_smooth_loss = F.smooth_l1_loss(rgb_full.reshape(-1, 3), pixels.reshape(-1, 3))
    loss = (rgb_discrete_loss + rgb_smooth_loss) / 2 + loss_reg
    optimizer.zero_grad(set_to_none=False)
    loss.sum().backward()
    if all_reduce is not None:
        all_reduce()
    optimizer.step()
    return dict(loss=loss.detach(), rgb_discrete_loss=rgb_discrete_loss.detach(), rgb_smooth_loss=rgb_smooth_loss.detach(),
                loss_reg=loss_reg.detach(), n_mesh_samples=n_mesh, n_volume_samples=n_vol)


def _flatten_rays(rays: Rays):
    rays_shape = rays.origins.shape
    if len(rays_shape) == 3:
        height, width, _ = rays_shape
        num_rays = height * width
        rays = namedtuple_map(lambda r: r.reshape([num_rays] + list(r.shape[2:])), rays)
    else:
        num_rays, _ = rays_shape
    return rays, rays_shape, num_rays


def _field_vector(field_net, x):
    """`del_vector[b:b+batch] = field_net(x, return_grad=False)[0]` (utils.py:557-565): the reference writes the
    (M, output_dim) output into an (M,3) buffer, so a scalar field is broadcast to the three components."""
    out = field_net(x, return_grad=False)[0]
    return out.expand(-1, 3) if out.shape[1] == 1 else out


def render_image_finetune_with_occgrid(radiance_field, field_net, estimator, rays: Rays, data, near_plane=0.0,
                                       far_plane=1e10, render_step_size=1e-3, render_bkgd=None, cone_angle=0.0,
                                       alpha_thre=0.0, test_chunk_size=8192, timestamps=None, mesh_intersect=None,
                                       mesh_finetune=None, scaling=1 / 128, bg_color="white"):
    """utils.py:465-607.  `data` is the reference tuple (xyzs, dirs, index_ray, ts, index_tri, origins).  Returns the
    9-tuple (rgb, opacity, depth, n_samples, weights, points, index_ray, loss, index_tri).

    Training mode (scaling != 0): the deformation field moves every quadrature point along its ray by
    tanh(field_net)·scaling (`dh`), the samples are re-sorted per ray, the radiance field is queried at the moved points
    (gradients reach `field_net` through the radiance field's POSITION gradient) and `mesh_finetune.update_d` accumulates
    the weighted displacement per triangle.  Reference quirks kept: `update_d` pairs the re-sorted `weights` with the
    un-sorted `dh` / `index_tri` (:588), and the vertex regulariser samples random points of the hit triangles (:543-546).
    With scaling == 0 (the evaluation passes, train_finetune.py:726) the displacement is exactly zero (quirk Q10) and
    `field_net` is not evaluated."""
    rays, rays_shape, num_rays = _flatten_rays(rays)
    dev = mesh_intersect.device
    xyzs, dirs, index_ray, ts, index_tri, origins = [t.to(dev) for t in data]
    index_ray, index_tri = index_ray.long(), index_tri.long()
    viewdirs = rays.viewdirs.to(dev)
    if scaling == 0:
        points, deltas, boundary, _, index_ray_s, depth, _, _ = mesh_intersect.sampling_indexing(xyzs, origins, dirs, index_ray, ts, index_tri)
        dh = torch.zeros_like(xyzs)
        loss = torch.zeros(1, device=dev)
    else:
        faces = torch.as_tensor(np.asarray(mesh_intersect.mesh.faces), device=dev).long()
        tri_v = mesh_intersect.vertices[faces[index_tri]][:, :, 0:3]                       # (M,3,3)   utils.py:543-544
        w = torch.rand((xyzs.shape[0], 3), device=xyzs.device)[..., None]
        vertices = torch.sum(tri_v * w, dim=1) / (torch.sum(w, dim=1) + 1e-6)
        del_vector_v = torch.tanh(_field_vector(field_net, vertices)) * scaling
        del_vector = torch.tanh(_field_vector(field_net, xyzs)) * scaling
        del_delta = (del_vector * dirs).sum(-1, keepdim=True)
        dh = del_delta * dirs
        xyzs = xyzs + dh
        ts = ts + del_delta.view(-1)
        # re-sorting the quadrature points based on the added dh
        points, deltas, boundary, _, index_ray_s, depth, _, _ = mesh_intersect.sampling_indexing(xyzs, origins, dirs, index_ray, ts, index_tri)
        loss = ((del_vector) ** 2).mean() + ((del_vector_v - del_vector.detach()) ** 2).mean()
        loss = loss.reshape(1)
    rgbs, sigmas = radiance_field(points, viewdirs, ray_indices=index_ray_s)                 # quirk Q7: original viewdirs
    rgb, opacity, _, depth_img, weights = derive_properties(rgbs, sigmas.squeeze(-1), depth, deltas, boundary, index_ray_s,
                                                            bg_color=bg_color, render_bkgd=render_bkgd, N=num_rays, return_ray_ids=False)
    if mesh_finetune is not None:
        mesh_finetune.update_d(dh, weights[:, 0], index_tri)
    return (rgb.view((*rays_shape[:-1], -1)), opacity.view((*rays_shape[:-1], -1)), depth_img.view((*rays_shape[:-1], -1)),
            xyzs.shape[0], weights, points, index_ray_s, loss, index_tri)


def render_image_fit_sg_with_occgrid(radiance_field, radiance_field_sg, estimator, rays: Rays, data, near_plane=0.0,
                                     far_plane=1e10, render_step_size=1e-3, render_bkgd=None, cone_angle=0.0,
                                     alpha_thre=0.0, test_chunk_size=8192, timestamps=None, mesh_intersect=None,
                                     mesh_finetune=None, scaling=1 / 128, bg_color="white"):
    """utils.py:610-731 (the render of train_fit_sg.py:439-452) -> the reference's 8-tuple
    (colors, opacities, depths, n_samples, weights, xyzs, index_ray, index_tri).  Colours come from the
    spherical-Gaussian field (with gradient), densities from the frozen radiance field (no gradient), both evaluated at the
    quadrature points with the ORIGINAL ray directions `rays.viewdirs[index_ray]` (:686, :663); constant quadrature step
    (quirk Q4, :709-710).  The reference's 32768-row loops are single kernel launches here."""
    rays, rays_shape, num_rays = _flatten_rays(rays)
    xyzs, dirs, index_ray, ts, index_tri, origins = data
    dev = radiance_field_sg.aabb.device
    xyzs, ts, index_ray = xyzs.to(dev), ts.to(dev), index_ray.to(dev).long()
    viewdirs = _lib.f32(rays.viewdirs, dev)
    rgbs, _ = radiance_field_sg(xyzs, viewdirs, ray_indices=index_ray)
    with torch.no_grad():
        _, sigmas = radiance_field(xyzs, viewdirs, ray_indices=index_ray)
        sigmas = sigmas.squeeze(-1)
        boundary = torch.ones_like(index_ray, dtype=torch.bool)
        boundary[1:] = index_ray[1:] != index_ray[:-1]                       # spc_render.mark_pack_boundaries (:708)
    rgb, opacity, _, depth, weights = derive_properties(rgbs, sigmas, ts, float(render_step_size), boundary, index_ray,
                                                        bg_color=bg_color, render_bkgd=render_bkgd, N=num_rays, return_ray_ids=False)
    return (rgb.view((*rays_shape[:-1], -1)), opacity.view((*rays_shape[:-1], -1)), depth.view((*rays_shape[:-1], -1)),
            xyzs.shape[0], weights, xyzs, index_ray, index_tri)


@torch.no_grad()
def render_image_bake_texture_images_with_occgrid(radiance_field, rays: Rays, data, texture=None, uv=None, near_plane=0.0,
                                                  far_plane=1e10, render_step_size=1e-3, render_bkgd=None, cone_angle=0.0,
                                                  alpha_thre=0.0, test_chunk_size=8192, timestamps=None,
                                                  mesh_intersect=None, mesh_finetune=None, scaling=1 / 128,
                                                  discretize=False, compressor=None, bg_color="white"):
    """utils.py:998-1095 -> the reference's 8-tuple.  Texel lookup (CPU fp64 trimesh barycentrics in the reference)
    and the 32 000-row decode loop are single kernels."""
    lib = _lib.load()
    rays, rays_shape, num_rays = _flatten_rays(rays)
    dev = mesh_intersect.device
    xyzs, dirs, index_ray, ts, index_tri, origins = [t.to(dev) for t in data]
    points, deltas, boundary, dirs, index_ray, depth, index_tri, _ = mesh_intersect.sampling_indexing(
        xyzs, origins, dirs, index_ray.long(), ts, index_tri.long())
    M = points.shape[0]
    uv_points = torch.empty((M, 2), dtype=torch.int64, device=dev)
    pts, tri, uvs = _lib.f32(points), _lib.i64(index_tri), _lib.f32(uv, dev)
    _lib.check(lib.qf_hit_texels(mesh_intersect.rayintersector.handle, _lib.ptr(pts), _lib.ptr(tri), M, _lib.ptr(uvs),
                                 compressor.texture_size, _lib.ptr(uv_points), _lib.stream(dev)), "qf_hit_texels")
    texture_points = compressor.get_features_from_texture_map(uv_points)
    if discretize:
        sigmas = inverse_of_compressed_sigma(compress_sigma(texture_points[:, -1]))
    else:
        sigmas = texture_points[:, -1]
    rgbs = radiance_field.features_to_rgb(texture_points[:, :-1], dirs)
    rgb, opacity, _, depth_img, weights = derive_properties(rgbs, sigmas, depth, deltas, boundary, index_ray,
                                                            bg_color=bg_color, render_bkgd=None, N=num_rays, return_ray_ids=False)
    return (rgb.view((*rays_shape[:-1], -1)), opacity.view((*rays_shape[:-1], -1)), depth_img.view((*rays_shape[:-1], -1)),
            xyzs.shape[0], weights, points, rays, 0)
